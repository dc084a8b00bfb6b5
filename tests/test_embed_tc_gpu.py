"""GPU: the tensor-core variant of the embedder's edge kernel (FCS_EMBED_MODE_TC: tcgen05, bf16 hi/lo split operands,
three products, fp32 accumulation in TMEM) against the same golden vectors and oracle, with the same tolerance as
the fp32 kernel.  Opt-in (FCS_TEST_EMBED_TC=1) until the mode has been validated on hardware and made the default."""
import os

import numpy as np
import pytest

from golden_util import embed_golden
from merizo_search_b200 import embed as b200_embed
from merizo_search_b200 import native
from oracle import foldclass_embed_oracle as emb

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FCS_TEST_EMBED_TC") != "1", reason="tensor-core embedder mode is opt-in")]
RTOL = 5e-5


@pytest.fixture(scope="module")
def golden():
    return embed_golden()


@pytest.fixture(scope="module")
def embedder(golden):
    e = b200_embed.FoldClassEmbedder(golden[1], device=0)
    e._emb.set_mode(native.EMBED_MODE_TC)
    yield e
    e.close()


def _rel(got, want):
    return float(np.abs(got - want).max() / np.abs(want).max())


def test_tc_layer_outputs(golden, embedder):
    z, sd, structures = golden
    c = structures[0]
    f0 = sd["posenc_as.pe"][0, :c.shape[0]]
    want_l0, want_m0 = emb.egnn_layer(f0, c, sd, 0, factored=True, return_messages=True)
    got_l0, got_m0 = embedder._emb.debug_layer(c, 0)
    errs = dict(m0=_rel(got_m0, want_m0), l0=_rel(got_l0, z["s0_layer0"]))
    print("tensor-core mode, relative errors per stage:", errs)
    if errs["m0"] > RTOL:  # localise: which (residue, channel) entries are off
        d = np.abs(got_m0 - want_m0) / np.abs(want_m0).max()
        bad = np.argwhere(d > RTOL)
        print("bad entries:", len(bad), "of", d.size, "first:", bad[:8].tolist(), "rows:", sorted(set(bad[:, 0].tolist()))[:16],
              "cols:", sorted(set(bad[:, 1].tolist()))[:16])
        print("got[0,:8]", got_m0[0, :8], "want[0,:8]", want_m0[0, :8])
    assert all(v < RTOL for v in errs.values()), errs


def test_tc_golden_and_ragged(golden, embedder):
    z, sd, structures = golden
    got = embedder.embed_structures(structures)
    bad = emb.embedding_close(got, z["embeddings"], rtol=RTOL)
    assert bad == [], bad
    rng = np.random.default_rng(12)
    lens = [int(x) for x in rng.integers(1, 200, size=16)] + [1, 16, 17, 300]
    chains = [emb.synthetic_chain(L, seed=700 + i) for i, L in enumerate(lens)]
    got = embedder.embed_structures(chains)
    bad = emb.embedding_close(got, emb.forward_batch(chains, sd, factored=True), rtol=RTOL)
    assert bad == [], bad


def test_tc_throughput_print(golden, embedder):
    from merizo_search_b200 import synth

    lens = synth.host_lengths(2048, seed=21)
    chains = synth.synthetic_chains(lens, seed=3)
    embedder.embed_structures(chains[:64])
    for mode, name in ((native.EMBED_MODE_TC, "tcgen05 bf16x3"), (native.EMBED_MODE_FP32, "fp32 FFMA2")):
        embedder._emb.set_mode(mode)
        embedder.embed_structures(chains)
        out = embedder.embed_structures(chains)
        t = embedder.timing()
        print(f"{name}: {len(chains)} structures {t.last_ms:.1f} ms (edge kernels {t.last_edge_ms:.1f} ms = "
              f"{2 * 2 * 514 * 256 * t.last_pairs / (t.last_edge_ms * 1e-3) / 1e12:.1f} algorithmic TFLOP/s)")
        if mode == native.EMBED_MODE_TC:
            tc_out = out
    assert emb.embedding_close(tc_out, out, rtol=RTOL) == []
    embedder._emb.set_mode(native.EMBED_MODE_TC)
