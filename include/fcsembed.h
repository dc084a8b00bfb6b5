/* fcsembed.h -- C ABI of the batched Foldclass query embedder (the step
 * immediately BEFORE the search hot path).
 *
 * The reference embeds one structure per call of a torch module and bounces every embedding
 * through the host (paths relative to merizo_search/programs/Foldclass/):
 *
 *   network(query_input)            dbsearch.py:97-98, dbsearch.py:287-301  -> fcs_embed / fcs_embed_to_device
 *   FoldClassNet.forward            nndef_fold_egnn_embed.py:50-62          -> the kernels behind them
 *   EGNN.forward                    my_egnn_nocoords.py:44-74               -> embed_edge_kernel (+ node kernels)
 *   network_setup / load_state_dict dbsearch.py:35-45                       -> fcs_embedder_create (weights by pointer)
 *
 * A whole ragged batch of C-alpha traces goes in (coords [sum L, 3] + offsets [n+1]); the [n,128]
 * embeddings come out either on the host or in device memory, ready for fcs_search_device -- the
 * queries never visit the host (the reference's .cpu() at dbsearch.py:316).
 *
 * Same conventions as fcsearch.h: plain pointers and sizes, FCS_* return codes, fcs_last_error().
 */
#ifndef FCSEMBED_H
#define FCSEMBED_H

#include <stdint.h>

#include "fcsearch.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FCS_EMBED_WIDTH 128   /* FoldClassNet(128): node feature width                      */
#define FCS_EMBED_HIDDEN 514  /* edge MLP hidden width = 2 * (2*128 + 1), my_egnn_nocoords.py:14-21 */
#define FCS_EMBED_MDIM 256    /* message width m_dim = 2 * 128, nndef_fold_egnn_embed.py:46  */
#define FCS_EMBED_MAX_LEN 3000 /* PositionalEncoder max_len, nndef_fold_egnn_embed.py:13     */

/* One EGNN layer's parameters, HOST pointers, fp32, nn.Linear layout ([out_features, in_features] row-major);
 * the state_dict keys are encode_ca_egnn.<layer>.<name>. */
typedef struct fcs_egnn_weights {
    const float* edge_w1; /* edge_mlp.0.weight  [514, 257]  (columns: feats_i 0..127, feats_j 128..255, dist^2 256) */
    const float* edge_b1; /* edge_mlp.0.bias    [514]       */
    const float* edge_w2; /* edge_mlp.2.weight  [256, 514]  */
    const float* edge_b2; /* edge_mlp.2.bias    [256]       */
    const float* gate_w;  /* edge_gate.0.weight [1, 256]    */
    const float* gate_b;  /* edge_gate.0.bias   [1]         */
    const float* node_w1; /* node_mlp.0.weight  [256, 384]  (columns: feats 0..127, summed messages 128..383) */
    const float* node_b1; /* node_mlp.0.bias    [256]       */
    const float* node_w2; /* node_mlp.2.weight  [128, 256]  */
    const float* node_b2; /* node_mlp.2.bias    [128]       */
} fcs_egnn_weights;

typedef struct fcs_embedder fcs_embedder;

typedef struct fcs_embed_timing {
    float last_ms;       /* device time of the last fcs_embed* call (CUDA events on the embedder's stream) */
    float last_edge_ms;  /* ... of its embed_edge_kernel launches (the dominant kernel) */
    int32_t last_launches;
    int32_t last_structures;
    int64_t last_residues;
    int64_t last_pairs;  /* sum over structures of L*L, per layer */
} fcs_embed_timing;

/* pe: the positional table posenc_as.pe, [max_len, 128] fp32 (host); max_len <= FCS_EMBED_MAX_LEN rows are kept. */
int fcs_embedder_create(int device, const fcs_egnn_weights* layers, int n_layers /* 1..4; FoldClassNet uses 2 */,
                        const float* pe, int max_len, fcs_embedder** out);
int fcs_embedder_destroy(fcs_embedder* e);

/* coords  [offsets[n], 3] fp32 HOST, structure s owns rows offsets[s] .. offsets[s+1]-1 (1 <= L <= max_len)
 * out     [n, 128] fp32 -- HOST memory for fcs_embed, DEVICE memory (the embedder's device) for
 *         fcs_embed_to_device.  Both return after the result is complete. */
int fcs_embed(fcs_embedder* e, const float* coords, const int64_t* offsets, int n_structures, float* out_host);
int fcs_embed_to_device(fcs_embedder* e, const float* coords, const int64_t* offsets, int n_structures, float* out_dev);

/* Which kernel evaluates the O(L^2) edge MLP:
 *   FCS_EMBED_MODE_TC3   tcgen05 tensor cores, bf16 hi/lo split operands, three products, fp32 accumulation in TMEM
 *                        (fp32-grade accuracy: the same parity tolerance applies); 16 operand-generator warps, dedicated
 *                        epilogue warps, two accumulator buffers (a tile's MMAs run under the previous tile's epilogue);
 *                        the default
 *   FCS_EMBED_MODE_TC2   same structure with 8 generator warps
 *   FCS_EMBED_MODE_TC    the round-1 tensor-core kernel (generators also run the epilogue, one accumulator buffer)
 *   FCS_EMBED_MODE_FP32  fp32 FMA pipe (packed FFMA2) */
#define FCS_EMBED_MODE_FP32 0
#define FCS_EMBED_MODE_TC 1
#define FCS_EMBED_MODE_TC2 2
#define FCS_EMBED_MODE_TC3 3
int fcs_embed_set_mode(fcs_embedder* e, int mode);

int fcs_embed_get_timing(const fcs_embedder* e, fcs_embed_timing* out);

/* Test hook: ONE structure; node features after EGNN layer `layer` ([L,128]) and that layer's summed
 * messages m_i ([L,256]) to HOST memory (either may be NULL). */
int fcs_embed_debug_layer(fcs_embedder* e, const float* coords, int length, int layer, float* out_feats, float* out_messages);

#ifdef __cplusplus
}
#endif
#endif /* FCSEMBED_H */
