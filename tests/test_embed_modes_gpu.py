"""GPU: the two edge kernels of the embedder agree with each other on a TED-like batch, and the mode switch is honoured
(prints the throughput of both for the record)."""
import numpy as np
import pytest

from golden_util import embed_golden
from merizo_search_b200 import embed as b200_embed
from merizo_search_b200 import native, synth
from oracle import foldclass_embed_oracle as emb

pytestmark = pytest.mark.gpu


def test_modes_agree_and_mode_errors():
    _, sd, _ = embed_golden()
    e = b200_embed.FoldClassEmbedder(sd, device=0)
    chains = synth.synthetic_chains(synth.host_lengths(512, seed=21), seed=3)
    outs = {}
    for mode, name in ((native.EMBED_MODE_TC3, "tcgen05 bf16x3, 16 generator warps (default)"), (native.EMBED_MODE_TC2, "tcgen05 bf16x3, 8 generator warps"),
                       (native.EMBED_MODE_TC, "tcgen05 bf16x3, round-1 kernel"), (native.EMBED_MODE_FP32, "fp32 FFMA2")):
        e._emb.set_mode(mode)
        e.embed_structures(chains[:32])
        outs[mode] = e.embed_structures(chains)
        t = e.timing()
        print(f"{name}: {len(chains)} structures {t.last_ms:.1f} ms (edge kernels {t.last_edge_ms:.1f} ms = "
              f"{2 * 2 * 514 * 256 * t.last_pairs / (t.last_edge_ms * 1e-3) / 1e12:.1f} algorithmic TFLOP/s)")
    for mode in (native.EMBED_MODE_TC3, native.EMBED_MODE_TC2, native.EMBED_MODE_TC):
        assert emb.embedding_close(outs[mode], outs[native.EMBED_MODE_FP32], rtol=5e-5) == [], mode
    assert not np.array_equal(outs[native.EMBED_MODE_TC3], outs[native.EMBED_MODE_FP32]), "mode switch had no effect"
    with pytest.raises(native.FcsError):
        e._emb.set_mode(7)
    e.close()
