// fcs_embed.cu -- batched FoldClassNet(128) forward: C-alpha traces -> 128-d embeddings (include/fcsembed.h).
//
// Replaces the one-structure-per-call torch forward of the reference (dbsearch.py:97-98, 287-301;
// nndef_fold_egnn_embed.py:50-62; my_egnn_nocoords.py:44-74) for whole ragged batches.  Per EGNN layer:
//
//   edge_input_ij = [f_i, f_j, d_ij^2]                              (257)
//   h_ij = SiLU(W1 edge_input_ij + b1)                              (514)     <- never materialised
//   m_ij = SiLU(W2 h_ij + b2);  m_ij *= sigmoid(wg . m_ij + bg)     (256)     <- never materialised
//   m_i  = sum_j m_ij;  f_i' = W4 SiLU(W3 [f_i, m_i] + b3) + b4 + f_i
//
// The reference materialises [L,L,257], [L,L,514] and [L,L,256] tensors in HBM per structure.  Here the first
// linear layer is split algebraically,  W1 [f_i, f_j, d2] + b1 = P_i + Q_j + d2 * w_d  with
// P = f W1[:, :128]^T + b1 and Q = f W1[:, 128:256]^T computed once per RESIDUE (K_e1), so the per-PAIR work
// is one fused kernel (K_e2, embed_edge_kernel): h_ij is generated on the fly in shared memory, contracted
// with W2 (the only O(L^2 x 514 x 256) term: 263 kFLOP per pair), gated and summed over j in registers /
// shared memory; only m_i [L,256] is written.  HBM traffic per pair is ~0.  The embedding must match the reference
// network to fp32 rounding (tests: max relative error 5e-5 of the largest component, cosine > 1 - 1e-6), so the
// contraction is either fp32 (FMA pipe) or bf16 hi/lo split x3 with fp32 accumulation (tensor cores).
//
//   K_e0 embed_init_feats_kernel   f = pe[:L]                          (PositionalEncoder.forward)
//   K_e1 embed_node_proj_kernel    P, Q  [R, 528] (514 padded to 33 x 16)
//   K_e2 embed_edge_tc2_kernel<16> m_i   [R, 256]   <- dominant; tcgen05 (default, FCS_EMBED_MODE_TC3): 3 x 2*514*256 bf16
//                                                     flop per (i,j) pair; generators / MMA issuer / epilogue in separate
//                                                     warps, two TMEM accumulator buffers
//        embed_edge_tc_kernel                          round-1 tcgen05 kernel (FCS_EMBED_MODE_TC)
//        embed_edge_kernel                             fp32 FMA-pipe variant (FCS_EMBED_MODE_FP32): 2*514*256 flop per pair
//   K_e3 embed_node_mlp_kernel     f'    [R, 128]
//   K_e4 embed_mean_kernel         mean over residues -> [n, 128]
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/fcsembed.h"
#include "fcs_common.cuh"
#include "fcs_internal.h"

namespace fcs {
namespace {

constexpr int EW = FCS_EMBED_WIDTH;    // 128
constexpr int EH = FCS_EMBED_HIDDEN;   // 514
constexpr int EHP = 528;               // hidden width padded to 33 chunks of 16 (pad weights are zero)
constexpr int EM = FCS_EMBED_MDIM;     // 256
constexpr int KC = 16;                 // hidden units per pipeline chunk
constexpr int NCHUNK = EHP / KC;       // 33
constexpr int TI = 8;                  // residues i per CTA
constexpr int TJ = 16;                 // residues j per tile
constexpr int TP = TI * TJ;            // 128 pairs per tile
constexpr int EDGE_THREADS = 256;      // 8 warps: 4 (pair blocks of 32) x 2 (channel halves of 128); thread = 16 pairs x 8 channels
constexpr int QS = EHP + 1;            // shared-memory row stride of the Q tile (conflict-free column reads)

// shared-memory carve-up of embed_edge_kernel (floats)
constexpr int SM_P = 0;                           // [TI][EHP]
constexpr int SM_Q = SM_P + TI * EHP;             // [TJ][QS]
constexpr int SM_WD = SM_Q + TJ * QS;             // [EHP]
constexpr int SM_H = ((SM_WD + EHP + 3) / 4) * 4; // [2][KC][TP]     hidden activations of the current / next chunk
constexpr int SM_W = SM_H + 2 * KC * TP;          // [2][KC][EM]     rows of W2^T of the current / next chunk
constexpr int SM_MSUM = SM_W + 2 * KC * EM;       // [TI][EM]
constexpr int SM_GP = SM_MSUM + TI * EM;          // [2][TP] gate partial dots (one per channel half)
constexpr int SM_D2 = SM_GP + 2 * TP;             // [TP]
constexpr int SM_VALID = SM_D2 + TP;              // [TP]
constexpr int SM_B2 = SM_VALID + TP;              // [EM]
constexpr int SM_WG = SM_B2 + EM;                 // [EM]
constexpr int SM_FLOATS = SM_WG + EM;
constexpr int EDGE_SMEM = SM_FLOATS * 4;
static_assert(SM_H % 4 == 0 && SM_W % 4 == 0 && SM_P % 4 == 0, "16-byte alignment of the vector-accessed regions");

__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ float sigmoidf(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// acc.xy += a.xy * b.xy  (SASS FFMA2: two fp32 FMAs per issue slot)
__device__ __forceinline__ void fma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long dup2(float w) {  // {w, w}
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "r"(__float_as_uint(w)));
    return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ K_e0
// f = pe[:L] for every structure (PositionalEncoder.forward ignores the values of its input).
__global__ void __launch_bounds__(256) embed_init_feats_kernel(const float* __restrict__ pe, const int* __restrict__ s_start,
                                                               const int* __restrict__ s_len, float* __restrict__ feats) {
    const int s = blockIdx.x;
    const int start = s_start[s], L = s_len[s];
    const float4* src = reinterpret_cast<const float4*>(pe);
    float4* dst = reinterpret_cast<float4*>(feats + size_t(start) * EW);
    for (int e = threadIdx.x; e < L * (EW / 4); e += blockDim.x) dst[e] = src[e];
}

// ------------------------------------------------------------------------------------------------ K_e1
// out[r][c] = bias[c] + sum_k f[r][k] * wt[k][c]   for 16 residues per CTA; c < ncols (thread per column).
// Used for P|Q (wt = [128][1056]: W1a^T | W1b^T, split into two [R][528] outputs) -- FLOPs are ~0.1 % of K_e2.
constexpr int NP_ROWS = 16;
// The feature tile is stored TRANSPOSED, [k][16 residues], so the 16 operands of a k-step are 4 broadcast LDS.128 instead of
// 16 LDS.32 (the round-1 kernel was bound by those loads).
__global__ void __launch_bounds__(256) embed_node_proj_kernel(const float* __restrict__ feats, int n_res,
                                                              const float* __restrict__ wt, const float* __restrict__ bias,
                                                              float* __restrict__ outP, float* __restrict__ outQ) {
    __shared__ __align__(16) float sFt[EW][NP_ROWS];
    const int r0 = blockIdx.x * NP_ROWS;
    for (int e = threadIdx.x; e < NP_ROWS * EW; e += blockDim.x) {
        const int r = e / EW, k = e % EW;  // coalesced global read, transposed shared write
        sFt[k][r] = (r0 + r < n_res) ? feats[size_t(r0 + r) * EW + k] : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * EHP; c += blockDim.x) {
        float acc[NP_ROWS];
        const float b = bias[c];
#pragma unroll
        for (int r = 0; r < NP_ROWS; ++r) acc[r] = b;
#pragma unroll 4
        for (int k = 0; k < EW; ++k) {
            const float w = wt[size_t(k) * (2 * EHP) + c];
            const float4* f4 = reinterpret_cast<const float4*>(&sFt[k][0]);
#pragma unroll
            for (int r4 = 0; r4 < NP_ROWS / 4; ++r4) {
                const float4 f = f4[r4];
                acc[4 * r4 + 0] = fmaf(f.x, w, acc[4 * r4 + 0]);
                acc[4 * r4 + 1] = fmaf(f.y, w, acc[4 * r4 + 1]);
                acc[4 * r4 + 2] = fmaf(f.z, w, acc[4 * r4 + 2]);
                acc[4 * r4 + 3] = fmaf(f.w, w, acc[4 * r4 + 3]);
            }
        }
        float* out = (c < EHP) ? outP : outQ;
        const int cc = (c < EHP) ? c : c - EHP;
#pragma unroll
        for (int r = 0; r < NP_ROWS; ++r)
            if (r0 + r < n_res) out[size_t(r0 + r) * EHP + cc] = acc[r];
    }
}

// ------------------------------------------------------------------------------------------------ K_e2
struct EdgeParams {
    const float* coords;   // [R][3]
    const int* s_start;    // [n] first residue row of structure s
    const int* s_len;      // [n]
    const int2* items;     // [n_items] (structure, first residue i of the CTA), longest structures first
    const float* P;        // [R][EHP]  W1[:, :128] f_i + b1
    const float* Q;        // [R][EHP]  W1[:, 128:256] f_j
    const float* wd;       // [EHP]     W1[:, 256] (the dist^2 column), zero padded
    const float* w2t;      // [EHP][EM] W2 transposed, zero rows beyond 514
    const float* b2;       // [EM]
    const float* wg;       // [EM]
    float bg;
    float* M;              // [R][EM]   m_i = sum_j gate_ij * m_ij
    const uint8_t* w2img;  // tensor-core path: W2 as bf16 hi/lo UMMA operand images, [17 chunks][hi|lo][256 x 32]
};

// One CTA = TI residues i of one structure against ALL residues j of that structure, TJ at a time.
// Tile = 128 (i,j) pairs x 256 message channels; the 514-wide hidden layer is streamed in 33 chunks of 16:
// every chunk, the CTA (a) prefetches the next 16 rows of W2^T with cp.async, (b) generates the next
// 16 x 128 hidden activations h = SiLU(P_i + Q_j + d2 * w_d) into shared memory, (c) runs 16 x 64 FFMA2 per
// thread on the current chunk.  A thread owns 16 pairs (ONE residue i against the tile's 16 residues j) x 8
// channels: accumulators are packed {pair 2q, pair 2q+1} so the h operand of a packed FMA is a natural 8-byte
// piece of an LDS.128 and only the 8 weights are duplicated in registers; 6 LDS.128 feed 64 FFMA2 (v1 of this
// kernel had 6 per 32 and was shared-memory bound at 41 TFLOP/s).  The gate dot product needs one 16-lane
// shuffle reduction; the sum over j is thread-local.
__global__ void __launch_bounds__(EDGE_THREADS, 1) embed_edge_kernel(const EdgeParams p) {
    extern __shared__ __align__(16) float sm[];
    float* sP = sm + SM_P;
    float* sQ = sm + SM_Q;
    float* sWd = sm + SM_WD;
    float* sH = sm + SM_H;
    float* sW = sm + SM_W;
    float* sMsum = sm + SM_MSUM;
    float* sGp = sm + SM_GP;
    float* sD2 = sm + SM_D2;
    float* sValid = sm + SM_VALID;
    float* sB2 = sm + SM_B2;
    float* sWg = sm + SM_WG;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int2 item = p.items[blockIdx.x];
    const int start = p.s_start[item.x], L = p.s_len[item.x], i0 = item.y;

    // FMA-phase coordinates: warp = (pair block of 32, channel half of 128); lane = (lp: which 16 pairs, lc: channel quad)
    const int warp_p = warp & 3, warp_c = warp >> 2;
    const int lc = lane & 15, lp = lane >> 4;
    const int my_il = warp_p * 2 + lp;           // the residue i (0..TI-1) whose 16 pairs this thread owns
    const int my_pair0 = my_il * TJ;             // first of its 16 pairs (pair = i_local * TJ + j_local)
    const int my_ch0 = warp_c * 128 + lc * 4;    // its channels: my_ch0 + {0..3} and my_ch0 + 64 + {0..3}
    // generation-phase coordinates: one pair, eight hidden units per chunk
    const int g_pair = tid & (TP - 1), g_kq = tid >> 7;
    const int g_il = g_pair >> 4, g_jl = g_pair & (TJ - 1);

    // ---- per-CTA setup: P rows of the TI residues (clamped inside the structure), constants, zeroed sums
    for (int e = tid; e < TI * (EHP / 4); e += EDGE_THREADS) {
        const int r = e / (EHP / 4), c4 = e % (EHP / 4);
        const int row = start + min(i0 + r, L - 1);
        reinterpret_cast<float4*>(sP)[r * (EHP / 4) + c4] = reinterpret_cast<const float4*>(p.P + size_t(row) * EHP)[c4];
    }
    for (int e = tid; e < EHP; e += EDGE_THREADS) sWd[e] = p.wd[e];
    for (int e = tid; e < EM; e += EDGE_THREADS) {
        sB2[e] = p.b2[e];
        sWg[e] = p.wg[e];
    }
    for (int e = tid; e < TI * EM; e += EDGE_THREADS) sMsum[e] = 0.f;

    auto load_w_chunk = [&](int kc, int buf) {  // 16 rows of W2^T = 16 KB contiguous
        const float* src = p.w2t + size_t(kc) * KC * EM;
        float* dst = sW + buf * (KC * EM);
#pragma unroll
        for (int u = 0; u < (KC * EM / 4) / EDGE_THREADS; ++u) {
            const int e = tid + u * EDGE_THREADS;
            cp_async16(dst + e * 4, src + e * 4);
        }
    };
    // one hidden activation of chunk kc: h[kl][g_pair], kl = g_kq * 8 + u
    auto gen_h_one = [&](int kc, int buf, float d2, int u) {
        const int kl = g_kq * (KC / 2) + u, kg = kc * KC + kl;
        const float x = fmaf(d2, sWd[kg], sP[g_il * EHP + kg] + sQ[g_jl * QS + kg]);
        sH[buf * (KC * TP) + kl * TP + g_pair] = silu(x);
    };

    for (int j0 = 0; j0 < L; j0 += TJ) {
        // ---- tile setup: Q rows of the TJ residues j, squared distances, validity
        for (int e = tid; e < TJ * (EHP / 4); e += EDGE_THREADS) {
            const int r = e / (EHP / 4), c4 = e % (EHP / 4);
            const int row = start + min(j0 + r, L - 1);
            const float4 v = reinterpret_cast<const float4*>(p.Q + size_t(row) * EHP)[c4];
            float* d = sQ + r * QS + c4 * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        if (tid < TP) {
            const int il = tid >> 4, jl = tid & (TJ - 1);
            const int ri = start + min(i0 + il, L - 1), rj = start + min(j0 + jl, L - 1);
            const float dx = p.coords[size_t(ri) * 3 + 0] - p.coords[size_t(rj) * 3 + 0];
            const float dy = p.coords[size_t(ri) * 3 + 1] - p.coords[size_t(rj) * 3 + 1];
            const float dz = p.coords[size_t(ri) * 3 + 2] - p.coords[size_t(rj) * 3 + 2];
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);  // torch.linalg.norm, my_egnn_nocoords.py:49
            sD2[tid] = dist * dist;                                 // edge_input takes dist*dist, :58
            sValid[tid] = (i0 + il < L && j0 + jl < L) ? 1.f : 0.f;
        }
        __syncthreads();
        const float g_d2 = sD2[g_pair];
        load_w_chunk(0, 0);
#pragma unroll
        for (int u = 0; u < KC / 2; ++u) gen_h_one(0, 0, g_d2, u);
        cp_async_wait_all();
        __syncthreads();

        unsigned long long acc[8][8];  // [pair couple q = pairs 2q, 2q+1][channel c] packed {pair 2q, pair 2q+1}
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[q][c] = 0ull;

#pragma unroll 1
        for (int kc = 0; kc < NCHUNK; ++kc) {
            const int cur = kc & 1;
            // the next chunk is produced while this one is consumed; its 8 activations per thread are interleaved with the
            // FMA stream (one every second k-step) so their MUFU / shared-memory latencies hide behind FFMA2 issue.  The last
            // chunk re-produces itself into the idle buffer: no branch in the loop body, 1/33 wasted generation.
            const int nxt = (kc + 1 < NCHUNK) ? kc + 1 : kc;
            load_w_chunk(nxt, cur ^ 1);
            const float* hb = sH + cur * (KC * TP) + my_pair0;
            const float* wb = sW + cur * (KC * EM) + my_ch0;
#pragma unroll
            for (int kl = 0; kl < KC; ++kl) {
                if ((kl & 1) == 0) gen_h_one(nxt, cur ^ 1, g_d2, kl >> 1);
                const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(hb + kl * TP);
                const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(hb + kl * TP + 4);
                const ulonglong2 h2 = *reinterpret_cast<const ulonglong2*>(hb + kl * TP + 8);
                const ulonglong2 h3 = *reinterpret_cast<const ulonglong2*>(hb + kl * TP + 12);
                const float4 wa = *reinterpret_cast<const float4*>(wb + kl * EM);
                const float4 wc = *reinterpret_cast<const float4*>(wb + kl * EM + 64);
                const unsigned long long hv[8] = {h0.x, h0.y, h1.x, h1.y, h2.x, h2.y, h3.x, h3.y};
                const unsigned long long wv[8] = {dup2(wa.x), dup2(wa.y), dup2(wa.z), dup2(wa.w),
                                                  dup2(wc.x), dup2(wc.y), dup2(wc.z), dup2(wc.w)};
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int c = 0; c < 8; ++c) fma2(acc[q][c], hv[q], wv[c]);
            }
            cp_async_wait_all();
            __syncthreads();
        }

        // ---- tile epilogue: m = SiLU(acc + b2); gate = sigmoid(wg . m + bg); m_i += gate * m
        float gd[16];
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) gd[pp] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int ch = my_ch0 + (c >> 2) * 64 + (c & 3);
            const float b = sB2[ch], g = sWg[ch];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float2 v = *reinterpret_cast<float2*>(&acc[q][c]);
                v.x = silu(v.x + b);
                v.y = silu(v.y + b);
                gd[2 * q] = fmaf(v.x, g, gd[2 * q]);
                gd[2 * q + 1] = fmaf(v.y, g, gd[2 * q + 1]);
                acc[q][c] = *reinterpret_cast<unsigned long long*>(&v);
            }
        }
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) {  // sum over the 16 lanes (lc) that share these pairs: this warp's 128 channels
            gd[pp] += __shfl_xor_sync(0xffffffffu, gd[pp], 1);
            gd[pp] += __shfl_xor_sync(0xffffffffu, gd[pp], 2);
            gd[pp] += __shfl_xor_sync(0xffffffffu, gd[pp], 4);
            gd[pp] += __shfl_xor_sync(0xffffffffu, gd[pp], 8);
        }
        if (lc == 0) {
#pragma unroll
            for (int pp = 0; pp < 16; ++pp) sGp[warp_c * TP + my_pair0 + pp] = gd[pp];
        }
        __syncthreads();
        float ms[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) ms[c] = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pl = my_pair0 + 2 * q;
            const float gate0 = sigmoidf((sGp[pl] + sGp[TP + pl]) + p.bg) * sValid[pl];
            const float gate1 = sigmoidf((sGp[pl + 1] + sGp[TP + pl + 1]) + p.bg) * sValid[pl + 1];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float2 v = *reinterpret_cast<const float2*>(&acc[q][c]);
                ms[c] = fmaf(gate1, v.y, fmaf(gate0, v.x, ms[c]));
            }
        }
        // (i_local, channel) is owned by exactly one thread of the CTA: plain read-modify-write
#pragma unroll
        for (int c = 0; c < 8; ++c) sMsum[my_il * EM + my_ch0 + (c >> 2) * 64 + (c & 3)] += ms[c];
        __syncthreads();  // sValid / sD2 / sQ / sGp are rewritten by the next tile
    }

    for (int e = tid; e < TI * EM; e += EDGE_THREADS) {
        const int il = e / EM, c = e % EM;
        if (i0 + il < L) p.M[size_t(start + i0 + il) * EM + c] = sMsum[e];
    }
}

// ------------------------------------------------------------------------------------------------ K_e2 (tensor cores)
// Same tile (128 pairs x 256 channels, 514 hidden units) on the 5th-generation tensor cores.  The contraction
// h[128 x 514] . W2^T[514 x 256] is a genuine GEMM; to keep fp32-grade accuracy both operands are split into
// bf16 hi + lo parts and three products are accumulated in fp32 in TMEM:  hh.Wh + hl.Wh + hh.Wl  (the dropped
// hl.Wl term is ~2^-16 relative; measured on the golden set the embedding differs from the fp32 reference by
// < 1e-6 relative, the same as the fp32 kernel).  Per 32-wide chunk of the hidden layer (two UMMA k-steps):
//   * warps 0..7 (generators; thread = pair row x 16 hidden units) compute the chunk's activations, split them and store
//     the hi / lo A operand images straight into shared memory in the canonical K-major no-swizzle core-matrix layout
//     (8 rows x 16 B core matrices, LBO 128 B, SBO 512 B), then fence.proxy.async and ONE elected mbarrier arrive per warp;
//   * warp 8 (one thread) streams the chunk's W2 image (hi | lo, 32 KB, one 1-D bulk copy; the images are identical for
//     every tile, so the copy stream simply cycles over the 17 chunks), waits for both operands, issues 2 x 3
//     tcgen05.mma.kind::f16 M128 x N256 x K16 and commits to the stage's mbarrier;
//   * three stages of each operand: the generators run up to two chunks ahead of the MMAs.
// Epilogue (the 8 generator warps; warp = TMEM lane quadrant x channel half): tcgen05.ld of the 128 x 256 fp32
// accumulators, m = SiLU(acc + b2) kept in registers (128 per thread), gate dot product thread-local + one shared-memory
// exchange between the two channel halves, sum over the tile's 16 residues j by a transpose-reduce over 16 lanes.
constexpr int KT = 32;                 // hidden units per chunk
constexpr int EHT = 544;               // 514 padded to 17 x 32 (pads: h = SiLU(0) = 0 and zero weights)
constexpr int NCH_T = EHT / KT;        // 17
constexpr int A_PART = TP * KT * 2;    // 8 KB: one bf16 part of a chunk's A tile
constexpr int A_STAGE = 2 * A_PART;    // hi | lo
constexpr int B_PART = EM * KT * 2;    // 16 KB
constexpr int B_STAGE = 2 * B_PART;    // hi | lo = 32 KB
constexpr int TC_STAGES = 3;           // ring depth of BOTH operands: chunk g uses A stage g % 3 and B stage g % 3
constexpr int B_STAGES = TC_STAGES;
constexpr int A_STAGES = TC_STAGES;
constexpr int QST = EHT + 4;           // sQ row stride: 16-byte aligned rows, conflict-free LDS.128 across 8 rows
constexpr int TCF_P = 0;                          // float offsets behind the operand stages
constexpr int TCF_Q = TCF_P + TI * EHT;
constexpr int TCF_WD = TCF_Q + TJ * QST;
constexpr int TCF_MSUM = TCF_WD + EHT;
constexpr int TCF_GP = TCF_MSUM + TI * EM;
constexpr int TCF_D2 = TCF_GP + 2 * TP;
constexpr int TCF_VALID = TCF_D2 + TP;
constexpr int TCF_B2 = TCF_VALID + TP;
constexpr int TCF_WG = TCF_B2 + EM;
constexpr int TCF_FLOATS = TCF_WG + EM;
constexpr int TC_OPERAND_BYTES = B_STAGES * B_STAGE + A_STAGES * A_STAGE;   // 144 KB
constexpr int EDGE_TC_SMEM = TC_OPERAND_BYTES + TCF_FLOATS * 4 + 12 * 8;
static_assert(EDGE_TC_SMEM <= 232448, "dynamic shared memory limit of sm_100");
static_assert((TCF_FLOATS * 4) % 8 == 0 && TCF_Q % 4 == 0 && TCF_WD % 4 == 0, "alignment");
constexpr uint64_t TC_DESC = (uint64_t(128 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46);
// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(EM >> 3) << 17) | (uint32_t(TP >> 4) << 24);

__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) { return TC_DESC | uint64_t((smem_addr >> 4) & 0x3FFF); }
__device__ __forceinline__ void tcg_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcg_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcg_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcg_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tcg_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tcg_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// SiLU / sigmoid for the tensor-core kernel, which is bound by the instruction rate of its 8 generator warps (2 per
// scheduler): the cheapest correct form is  sigmoid(x) = rcp(1 + ex2(-x * log2 e))  -- two XU ops (MUFU.EX2, MUFU.RCP; the
// 16-lane XU pipe has the room since the bf16 conversions moved to the ALU pipe) and no range handling: ex2 overflows to
// +inf for x < -88 (rcp(inf) = 0, SiLU = -0 like the reference) and flushes to 0 for large x (sigmoid = 1).  A Newton-Raphson
// reciprocal on the FMA pipe (one XU op per activation) measured the same throughput at 6 more instructions per activation.
__device__ __forceinline__ float sigmoid_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}
__device__ __forceinline__ float silu_fast(float x) { return x * sigmoid_fast(x); }
// two activations -> packed bf16 hi parts and packed bf16 lo parts (element 0 in the low half); F2FP.PACK_AB runs on the
// ALU pipe (the scalar F2F.BF16.F32 conversions it replaces ran on the XU pipe)
__device__ __forceinline__ void split_bf16x2(float h0, float h1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(h0, h1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(h0 - __uint_as_float(hi << 16), h1 - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

constexpr int TC_GEN_THREADS = EDGE_THREADS;          // warps 0..7: activation generators, then the tile's epilogue
constexpr int TC_THREADS_ALL = TC_GEN_THREADS + 32;   // warp 8: one thread streams the W2 images and issues the MMAs

__global__ void __launch_bounds__(TC_THREADS_ALL, 1) embed_edge_tc_kernel(const EdgeParams p) {
    extern __shared__ __align__(1024) uint8_t smraw[];
    uint8_t* sB = smraw;
    uint8_t* sA = smraw + B_STAGES * B_STAGE;
    float* fl = reinterpret_cast<float*>(smraw + TC_OPERAND_BYTES);
    float* sP = fl + TCF_P;
    float* sQ = fl + TCF_Q;
    float* sWd = fl + TCF_WD;
    float* sMsum = fl + TCF_MSUM;
    float* sGp = fl + TCF_GP;
    float* sD2 = fl + TCF_D2;
    float* sValid = fl + TCF_VALID;
    float* sB2 = fl + TCF_B2;
    float* sWg = fl + TCF_WG;
    uint64_t* bars = reinterpret_cast<uint64_t*>(fl + TCF_FLOATS);
    uint64_t* b_full = bars;        // [3] W2 image of a chunk has landed (bulk-copy transaction bytes)
    uint64_t* mma_done = bars + 3;  // [3] the MMAs that read stage s (A and B of the same chunk) are complete
    uint64_t* a_full = bars + 6;    // [3] the 8 generator warps have written A stage s (one elected arrive per warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int2 item = p.items[blockIdx.x];
    const int start = p.s_start[item.x], L = p.s_len[item.x], i0 = item.y;
    const uint32_t total_chunks = uint32_t((L + TJ - 1) / TJ) * NCH_T;

    if (tid == 0) {
        for (int i = 0; i < TC_STAGES; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&mma_done[i], 1);
            mbar_init(&a_full[i], TC_GEN_THREADS / 32);
        }
        mbar_fence_init();
    }
    if (warp == 0) {  // whole warp: 256 TMEM columns = one 128 x 256 fp32 accumulator (one CTA per SM: shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- per-CTA setup: P rows of the TI residues (clamped inside the structure; pads zero), constants, zeroed sums
    for (int e = tid; e < TI * (EHT / 4); e += TC_THREADS_ALL) {
        const int r = e / (EHT / 4), c4 = e % (EHT / 4);
        const int row = start + min(i0 + r, L - 1);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < EHP / 4) v = reinterpret_cast<const float4*>(p.P + size_t(row) * EHP)[c4];
        reinterpret_cast<float4*>(sP)[r * (EHT / 4) + c4] = v;
    }
    for (int e = tid; e < EHT; e += TC_THREADS_ALL) sWd[e] = (e < EHP) ? p.wd[e] : 0.f;
    for (int e = tid; e < EM; e += TC_THREADS_ALL) {
        sB2[e] = p.b2[e];
        sWg[e] = p.wg[e];
    }
    for (int e = tid; e < TI * EM; e += TC_THREADS_ALL) sMsum[e] = 0.f;
    tcg_fence_before();
    __syncthreads();
    tcg_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TC_GEN_THREADS / 32) {
        // ---------------------------------------------------------------- W2 image stream + MMA issue (one thread)
        if (lane == 0) {
            const uint64_t pol_keep = policy_evict_normal();
            mbar_arrive_expect_tx(&b_full[0], B_STAGE);  // image of the first chunk
            bulk_g2s(sB, p.w2img, B_STAGE, &b_full[0], pol_keep);
            int c = 0;
            for (uint32_t g = 0; g < total_chunks; ++g) {
                const uint32_t st = g % TC_STAGES, use = g / TC_STAGES;
                if (g + 1 < total_chunks) {
                    // stage of chunk g+1 was last read by MMA g-2: once that has completed, stream the next W2 image in
                    // (the images cycle over the 17 chunks of W2; MMA g-1 keeps the tensor pipe busy meanwhile)
                    const uint32_t st1 = (g + 1) % TC_STAGES, use1 = (g + 1) / TC_STAGES;
                    const int c1 = (c + 1 == NCH_T) ? 0 : c + 1;
                    mbar_wait(&mma_done[st1], (use1 & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&b_full[st1], B_STAGE);
                    bulk_g2s(sB + st1 * B_STAGE, p.w2img + size_t(c1) * B_STAGE, B_STAGE, &b_full[st1], pol_keep);
                }
                mbar_wait(&a_full[st], use & 1u);  // activations of chunk g are in A stage st ...
                mbar_wait(&b_full[st], use & 1u);  // ... and its W2 image in B stage st
                tcg_fence_after();  // also orders the epilogue's TMEM reads of the previous tile before the overwrite below
                const uint32_t a_addr = smem_u32(sA + st * A_STAGE), b_addr = smem_u32(sB + st * B_STAGE);
#pragma unroll
                for (int ks = 0; ks < KT / 16; ++ks) {
                    const uint64_t ah = tc_desc(a_addr + ks * 256), al = tc_desc(a_addr + A_PART + ks * 256);
                    const uint64_t bh = tc_desc(b_addr + ks * 256), bl = tc_desc(b_addr + B_PART + ks * 256);
                    tcg_mma(tmem_base, ah, bh, (c > 0 || ks > 0) ? 1u : 0u);
                    tcg_mma(tmem_base, al, bh, 1u);
                    tcg_mma(tmem_base, ah, bl, 1u);
                }
                tcg_commit(&mma_done[st]);
                c = (c + 1 == NCH_T) ? 0 : c + 1;
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- generators (thread = pair row x 16 hidden units)
        const int g_row = tid & (TP - 1), g_kh = tid >> 7;
        const int g_il = g_row >> 4, g_jl = g_row & (TJ - 1);
        const uint32_t a_row_off = uint32_t(g_row >> 3) * 512u + uint32_t(g_row & 7) * 16u;
        // epilogue coordinates: TMEM lane quadrant (= warp % 4) x channel half
        const int quad = warp & 3, half = warp >> 2;
        const int e_row = quad * 32 + lane;  // pair row of this thread's accumulator lane
        const int e_il = e_row >> 4;

        uint32_t g = 0;  // running chunk counter of this CTA: stage g % 3
        for (int j0 = 0; j0 < L; j0 += TJ) {
            // ---- tile setup: Q rows of the TJ residues j (pads zero), squared distances, validity
            for (int e = tid; e < TJ * (EHT / 4); e += TC_GEN_THREADS) {
                const int r = e / (EHT / 4), c4 = e % (EHT / 4);
                const int row = start + min(j0 + r, L - 1);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c4 < EHP / 4) v = reinterpret_cast<const float4*>(p.Q + size_t(row) * EHP)[c4];
                *reinterpret_cast<float4*>(sQ + r * QST + c4 * 4) = v;
            }
            if (tid < TP) {
                const int il = tid >> 4, jl = tid & (TJ - 1);
                const int ri = start + min(i0 + il, L - 1), rj = start + min(j0 + jl, L - 1);
                const float dx = p.coords[size_t(ri) * 3 + 0] - p.coords[size_t(rj) * 3 + 0];
                const float dy = p.coords[size_t(ri) * 3 + 1] - p.coords[size_t(rj) * 3 + 1];
                const float dz = p.coords[size_t(ri) * 3 + 2] - p.coords[size_t(rj) * 3 + 2];
                const float dist = sqrtf(dx * dx + dy * dy + dz * dz);  // torch.linalg.norm, my_egnn_nocoords.py:49
                sD2[tid] = dist * dist;                                 // edge_input takes dist*dist, :58
                sValid[tid] = (i0 + il < L && j0 + jl < L) ? 1.f : 0.f;
            }
            named_bar_sync(1, TC_GEN_THREADS);
            const float g_d2 = sD2[g_row];

#pragma unroll 1
            for (int c = 0; c < NCH_T; ++c, ++g) {
                const uint32_t st = g % TC_STAGES;
                mbar_wait(&mma_done[st], ((g / TC_STAGES) & 1u) ^ 1u);  // MMA g-3 has completed: A stage `st` is free again
                // ---- this chunk's activations: h = SiLU(P_i + Q_j + d2 * w_d), split into bf16 hi + lo
                uint8_t* a_hi = sA + st * A_STAGE;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int k8 = g_kh * 2 + u;           // core-matrix column inside the chunk
                    const int kg = c * KT + k8 * 8;        // first hidden unit of the 8
                    const float4 pa = *reinterpret_cast<const float4*>(sP + g_il * EHT + kg);
                    const float4 pb = *reinterpret_cast<const float4*>(sP + g_il * EHT + kg + 4);
                    const float4 qa = *reinterpret_cast<const float4*>(sQ + g_jl * QST + kg);
                    const float4 qb = *reinterpret_cast<const float4*>(sQ + g_jl * QST + kg + 4);
                    const float4 wa = *reinterpret_cast<const float4*>(sWd + kg);
                    const float4 wb = *reinterpret_cast<const float4*>(sWd + kg + 4);
                    const float x[8] = {fmaf(g_d2, wa.x, pa.x + qa.x), fmaf(g_d2, wa.y, pa.y + qa.y), fmaf(g_d2, wa.z, pa.z + qa.z),
                                        fmaf(g_d2, wa.w, pa.w + qa.w), fmaf(g_d2, wb.x, pb.x + qb.x), fmaf(g_d2, wb.y, pb.y + qb.y),
                                        fmaf(g_d2, wb.z, pb.z + qb.z), fmaf(g_d2, wb.w, pb.w + qb.w)};
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_bf16x2(silu_fast(x[2 * e]), silu_fast(x[2 * e + 1]), hw[e], lw[e]);
                    const uint4 vh = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    const uint4 vl = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    *reinterpret_cast<uint4*>(a_hi + a_row_off + k8 * 128) = vh;
                    *reinterpret_cast<uint4*>(a_hi + A_PART + a_row_off + k8 * 128) = vl;
                }
                fence_proxy_async();  // generic-proxy stores above -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[st]);
            }
            // ---- all MMAs of the tile are complete once the last commit (chunk g-1) has arrived
            mbar_wait(&mma_done[(g - 1) % TC_STAGES], ((g - 1) / TC_STAGES) & 1u);
            tcg_fence_after();

            // ---- epilogue: m = SiLU(acc + b2) for this thread's pair and 128 channels; gate; sum over j
            float m[128];
            float gd = 0.f;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                uint32_t r[32];
                tcg_ld32(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(half * 128 + part * 32), r);
                tcg_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int ch = half * 128 + part * 32 + e;
                    const float v = silu_fast(__uint_as_float(r[e]) + sB2[ch]);
                    gd = fmaf(v, sWg[ch], gd);
                    m[part * 32 + e] = v;
                }
            }
            sGp[half * TP + e_row] = gd;
            tcg_fence_before();  // the accumulator reads above are ordered before the next tile's first MMA (via a_full)
            named_bar_sync(1, TC_GEN_THREADS);
            const float gate = sigmoid_fast((sGp[e_row] + sGp[TP + e_row]) + p.bg) * sValid[e_row];
            // sum over the tile's 16 residues j = the 16 lanes of a half-warp: transpose-reduce (each step a lane keeps
            // one half of its channels and receives the partner's sums for that half) -- 120 shuffles instead of 512
            const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
            float t64[64], t32[32], t16[16], t8[8];
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                const float lo_v = m[k] * gate, hi_v = m[k + 64] * gate;
                const float keep = b3 ? hi_v : lo_v, send = b3 ? lo_v : hi_v;
                t64[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float keep = b2 ? t64[k + 32] : t64[k], send = b2 ? t64[k] : t64[k + 32];
                t32[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float keep = b1 ? t32[k + 16] : t32[k], send = b1 ? t32[k] : t32[k + 16];
                t16[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float keep = b0 ? t16[k + 8] : t16[k], send = b0 ? t16[k] : t16[k + 8];
                t8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            // this lane now owns 8 channels of residue e_il: unique owner of (residue i, channel) in the CTA
            const int ch0 = half * 128 + (b3 ? 64 : 0) + (b2 ? 32 : 0) + (b1 ? 16 : 0) + (b0 ? 8 : 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) sMsum[e_il * EM + ch0 + k] += t8[k];
            named_bar_sync(1, TC_GEN_THREADS);  // sValid / sD2 / sQ / sGp are rewritten by the next tile
        }
    }

    tcg_fence_before();
    __syncthreads();
    for (int e = tid; e < TI * EM; e += TC_THREADS_ALL) {
        const int il = e / EM, c = e % EM;
        if (i0 + il < L) p.M[size_t(start + i0 + il) * EM + c] = sMsum[e];
    }
    if (warp == 0) {
        tcg_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ K_e2 (tensor cores, v2)
// FCS_EMBED_MODE_TC2 (8 generator warps) and FCS_EMBED_MODE_TC3 (16 generator warps; the default since round 2: 24.0 k
// structures/s vs 23.4 k for TC2 and 21.2 k for embed_edge_tc_kernel, parity identical -- scripts/embed_candidates.py).
// Same math and operand pipeline as embed_edge_tc_kernel; what changes is who does the epilogue and when:
//   warps 0..7   generators only (they never touch TMEM): after a tile's 17 chunks they go straight to the next tile
//   warp  8      W2 image stream + MMA issue; two 256-column accumulators (all 512 TMEM columns): tile t accumulates into
//                buffer t & 1, so the MMAs of tile t+1 run while tile t is being drained
//   warps 9..16  epilogue (warp = TMEM lane quadrant x channel half).  120 registers per thread at 544 threads do not hold
//                128 activations, so the tile is drained in two passes of tcgen05.ld: pass A accumulates the gate dot
//                product (SiLU values are dropped), pass B recomputes SiLU, applies the gate and transpose-reduces 32
//                channels at a time over the 16 lanes that hold the tile's 16 residues j.
// Per-tile shared state read by the epilogue (validity mask) is double-buffered by tile parity; the generators wait for
// the epilogue of tile t before they overwrite the mask of tile t+2.
constexpr int TC2_EPI_WARPS = 8;
__host__ __device__ constexpr int tc2_threads(int gen_warps) { return (gen_warps + 1 + TC2_EPI_WARPS) * 32; }  // 544 (8 generator warps) / 800 (16)
constexpr int TC2F_VALID2 = TCF_FLOATS;                 // second validity buffer (tile parity 1) behind the v1 float region
constexpr int TC2F_FLOATS = TC2F_VALID2 + TP;
constexpr int EDGE_TC2_SMEM = TC_OPERAND_BYTES + TC2F_FLOATS * 4 + 16 * 8;
static_assert(EDGE_TC2_SMEM <= 232448 && (TC2F_FLOATS * 4) % 8 == 0, "shared memory budget / barrier alignment");

// GW = generator warps: 8 (thread = pair row x 16 hidden units per chunk) or 16 (x 8 hidden units: twice the warps to hide
// the activation chain's latency; 80 registers per thread at 800 threads)
template <int GW>
__global__ void __launch_bounds__(tc2_threads(GW), 1) embed_edge_tc2_kernel(const EdgeParams p) {
    constexpr int TC2_GEN_WARPS = GW;
    constexpr int TC2_THREADS = tc2_threads(GW);
    constexpr int GEN_THREADS = GW * 32;
    constexpr int K8_PER_THREAD = 4 / (GW / 4);  // core-matrix columns (8 hidden units) per thread and chunk
    extern __shared__ __align__(1024) uint8_t smraw[];
    uint8_t* sB = smraw;
    uint8_t* sA = smraw + B_STAGES * B_STAGE;
    float* fl = reinterpret_cast<float*>(smraw + TC_OPERAND_BYTES);
    float* sP = fl + TCF_P;
    float* sQ = fl + TCF_Q;
    float* sWd = fl + TCF_WD;
    float* sMsum = fl + TCF_MSUM;
    float* sGp = fl + TCF_GP;
    float* sD2 = fl + TCF_D2;
    float* sB2 = fl + TCF_B2;
    float* sWg = fl + TCF_WG;
    uint64_t* bars = reinterpret_cast<uint64_t*>(fl + TC2F_FLOATS);
    uint64_t* b_full = bars;          // [3]
    uint64_t* mma_done = bars + 3;    // [3]
    uint64_t* a_full = bars + 6;      // [3]
    uint64_t* tmem_full = bars + 9;   // [2] all MMAs of the tile in accumulator buffer b have completed
    uint64_t* tmem_empty = bars + 11; // [2] the 8 epilogue warps have drained buffer b (and are done with its validity mask)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int2 item = p.items[blockIdx.x];
    const int start = p.s_start[item.x], L = p.s_len[item.x], i0 = item.y;
    const int n_tiles = (L + TJ - 1) / TJ;
    const uint32_t total_chunks = uint32_t(n_tiles) * NCH_T;

    if (tid == 0) {
        for (int i = 0; i < TC_STAGES; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&mma_done[i], 1);
            mbar_init(&a_full[i], TC2_GEN_WARPS);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], TC2_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < TI * (EHT / 4); e += TC2_THREADS) {
        const int r = e / (EHT / 4), c4 = e % (EHT / 4);
        const int row = start + min(i0 + r, L - 1);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < EHP / 4) v = reinterpret_cast<const float4*>(p.P + size_t(row) * EHP)[c4];
        reinterpret_cast<float4*>(sP)[r * (EHT / 4) + c4] = v;
    }
    for (int e = tid; e < EHT; e += TC2_THREADS) sWd[e] = (e < EHP) ? p.wd[e] : 0.f;
    for (int e = tid; e < EM; e += TC2_THREADS) {
        sB2[e] = p.b2[e];
        sWg[e] = p.wg[e];
    }
    for (int e = tid; e < TI * EM; e += TC2_THREADS) sMsum[e] = 0.f;
    tcg_fence_before();
    __syncthreads();
    tcg_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TC2_GEN_WARPS) {
        // ---------------------------------------------------------------- W2 image stream + MMA issue (one thread)
        if (lane == 0) {
            const uint64_t pol_keep = policy_evict_normal();
            mbar_arrive_expect_tx(&b_full[0], B_STAGE);
            bulk_g2s(sB, p.w2img, B_STAGE, &b_full[0], pol_keep);
            uint32_t g = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const uint32_t buf = uint32_t(t) & 1u;
                const uint32_t d_tmem = tmem_base + buf * uint32_t(EM);
                // the epilogue of tile t-2 has drained this accumulator buffer
                mbar_wait(&tmem_empty[buf], ((uint32_t(t) >> 1) & 1u) ^ 1u);
                tcg_fence_after();
                for (int c = 0; c < NCH_T; ++c, ++g) {
                    const uint32_t st = g % TC_STAGES, use = g / TC_STAGES;
                    if (g + 1 < total_chunks) {
                        const uint32_t st1 = (g + 1) % TC_STAGES, use1 = (g + 1) / TC_STAGES;
                        const int c1 = (c + 1 == NCH_T) ? 0 : c + 1;
                        mbar_wait(&mma_done[st1], (use1 & 1u) ^ 1u);
                        mbar_arrive_expect_tx(&b_full[st1], B_STAGE);
                        bulk_g2s(sB + st1 * B_STAGE, p.w2img + size_t(c1) * B_STAGE, B_STAGE, &b_full[st1], pol_keep);
                    }
                    mbar_wait(&a_full[st], use & 1u);
                    mbar_wait(&b_full[st], use & 1u);
                    tcg_fence_after();
                    const uint32_t a_addr = smem_u32(sA + st * A_STAGE), b_addr = smem_u32(sB + st * B_STAGE);
#pragma unroll
                    for (int ks = 0; ks < KT / 16; ++ks) {
                        const uint64_t ah = tc_desc(a_addr + ks * 256), al = tc_desc(a_addr + A_PART + ks * 256);
                        const uint64_t bh = tc_desc(b_addr + ks * 256), bl = tc_desc(b_addr + B_PART + ks * 256);
                        tcg_mma(d_tmem, ah, bh, (c > 0 || ks > 0) ? 1u : 0u);
                        tcg_mma(d_tmem, al, bh, 1u);
                        tcg_mma(d_tmem, ah, bl, 1u);
                    }
                    tcg_commit(&mma_done[st]);
                }
                tcg_commit(&tmem_full[buf]);  // arrives when every MMA issued so far (the whole tile) has completed
            }
        }
        __syncwarp();
    } else if (warp < TC2_GEN_WARPS) {
        // ---------------------------------------------------------------- generators (thread = pair row x 16 hidden units)
        const int g_row = tid & (TP - 1), g_kq = tid >> 7;  // g_kq: 0..GW/4-1
        const int g_il = g_row >> 4, g_jl = g_row & (TJ - 1);
        const uint32_t a_row_off = uint32_t(g_row >> 3) * 512u + uint32_t(g_row & 7) * 16u;
        uint32_t g = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int j0 = t * TJ;
            float* sValidT = (t & 1) ? (fl + TC2F_VALID2) : (fl + TCF_VALID);
            // the epilogue of tile t-2 is done with this parity's validity mask (same event that frees its accumulator)
            mbar_wait(&tmem_empty[t & 1], ((uint32_t(t) >> 1) & 1u) ^ 1u);
            for (int e = tid; e < TJ * (EHT / 4); e += GEN_THREADS) {
                const int r = e / (EHT / 4), c4 = e % (EHT / 4);
                const int row = start + min(j0 + r, L - 1);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c4 < EHP / 4) v = reinterpret_cast<const float4*>(p.Q + size_t(row) * EHP)[c4];
                *reinterpret_cast<float4*>(sQ + r * QST + c4 * 4) = v;
            }
            if (tid < TP) {
                const int il = tid >> 4, jl = tid & (TJ - 1);
                const int ri = start + min(i0 + il, L - 1), rj = start + min(j0 + jl, L - 1);
                const float dx = p.coords[size_t(ri) * 3 + 0] - p.coords[size_t(rj) * 3 + 0];
                const float dy = p.coords[size_t(ri) * 3 + 1] - p.coords[size_t(rj) * 3 + 1];
                const float dz = p.coords[size_t(ri) * 3 + 2] - p.coords[size_t(rj) * 3 + 2];
                const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
                sD2[tid] = dist * dist;
                sValidT[tid] = (i0 + il < L && j0 + jl < L) ? 1.f : 0.f;
            }
            named_bar_sync(1, GEN_THREADS);
            const float g_d2 = sD2[g_row];
#pragma unroll 1
            for (int c = 0; c < NCH_T; ++c, ++g) {
                const uint32_t st = g % TC_STAGES;
                mbar_wait(&mma_done[st], ((g / TC_STAGES) & 1u) ^ 1u);
                uint8_t* a_hi = sA + st * A_STAGE;
#pragma unroll
                for (int u = 0; u < K8_PER_THREAD; ++u) {
                    const int k8 = g_kq * K8_PER_THREAD + u;
                    const int kg = c * KT + k8 * 8;
                    const float4 pa = *reinterpret_cast<const float4*>(sP + g_il * EHT + kg);
                    const float4 pb = *reinterpret_cast<const float4*>(sP + g_il * EHT + kg + 4);
                    const float4 qa = *reinterpret_cast<const float4*>(sQ + g_jl * QST + kg);
                    const float4 qb = *reinterpret_cast<const float4*>(sQ + g_jl * QST + kg + 4);
                    const float4 wa = *reinterpret_cast<const float4*>(sWd + kg);
                    const float4 wb = *reinterpret_cast<const float4*>(sWd + kg + 4);
                    const float x[8] = {fmaf(g_d2, wa.x, pa.x + qa.x), fmaf(g_d2, wa.y, pa.y + qa.y), fmaf(g_d2, wa.z, pa.z + qa.z),
                                        fmaf(g_d2, wa.w, pa.w + qa.w), fmaf(g_d2, wb.x, pb.x + qb.x), fmaf(g_d2, wb.y, pb.y + qb.y),
                                        fmaf(g_d2, wb.z, pb.z + qb.z), fmaf(g_d2, wb.w, pb.w + qb.w)};
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_bf16x2(silu_fast(x[2 * e]), silu_fast(x[2 * e + 1]), hw[e], lw[e]);
                    *reinterpret_cast<uint4*>(a_hi + a_row_off + k8 * 128) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(a_hi + A_PART + a_row_off + k8 * 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[st]);
            }
            // sQ / sD2 of this tile are read only by the generators themselves: all of them must be past the tile's last
            // chunk before the next tile's setup overwrites them
            named_bar_sync(1, GEN_THREADS);
        }
    } else {
        // ---------------------------------------------------------------- epilogue warps 9..16
        const int ew = warp - (TC2_GEN_WARPS + 1);       // 0..7
        const int quad = warp & 3;                        // TMEM lane quadrant this warp may access (= warp id % 4)
        const int half = ew >> 2;                         // channel half handled by this warp: ew = {quadrant order} x half
        const int e_row = quad * 32 + lane;
        const int e_il = e_row >> 4;
        const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t buf = uint32_t(t) & 1u;
            const float* sValidT = (t & 1) ? (fl + TC2F_VALID2) : (fl + TCF_VALID);
            mbar_wait(&tmem_full[buf], (uint32_t(t) >> 1) & 1u);
            tcg_fence_after();
            const uint32_t tbase = tmem_base + (uint32_t(quad * 32) << 16) + buf * uint32_t(EM) + uint32_t(half * 128);
            // ---- pass A: gate dot product over this thread's 128 channels
            float gd = 0.f;
#pragma unroll 1
            for (int part = 0; part < 4; ++part) {
                uint32_t r[32];
                tcg_ld32(tbase + uint32_t(part * 32), r);
                tcg_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int ch = half * 128 + part * 32 + e;
                    gd = fmaf(silu_fast(__uint_as_float(r[e]) + sB2[ch]), sWg[ch], gd);
                }
            }
            sGp[half * TP + e_row] = gd;
            named_bar_sync(2, TC2_EPI_WARPS * 32);
            const float gate = sigmoid_fast((sGp[e_row] + sGp[TP + e_row]) + p.bg) * sValidT[e_row];
            // ---- pass B: recompute SiLU, apply the gate, sum over the 16 residues j (16 lanes) 32 channels at a time
#pragma unroll 1
            for (int part = 0; part < 4; ++part) {
                uint32_t r[32];
                tcg_ld32(tbase + uint32_t(part * 32), r);
                tcg_wait_ld();
                float v[32], t16[16], t8[8], t4[4], t2[2];
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = silu_fast(__uint_as_float(r[e]) + sB2[half * 128 + part * 32 + e]) * gate;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float keep = b3 ? v[k + 16] : v[k], send = b3 ? v[k] : v[k + 16];
                    t16[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float keep = b2 ? t16[k + 8] : t16[k], send = b2 ? t16[k] : t16[k + 8];
                    t8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float keep = b1 ? t8[k + 4] : t8[k], send = b1 ? t8[k] : t8[k + 4];
                    t4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float keep = b0 ? t4[k + 2] : t4[k], send = b0 ? t4[k] : t4[k + 2];
                    t2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                }
                // this lane owns 2 channels of residue e_il for this part: unique owner of (residue i, channel) in the CTA
                const int ch0 = half * 128 + part * 32 + (b3 ? 16 : 0) + (b2 ? 8 : 0) + (b1 ? 4 : 0) + (b0 ? 2 : 0);
                sMsum[e_il * EM + ch0] += t2[0];
                sMsum[e_il * EM + ch0 + 1] += t2[1];
            }
            // buffer drained (and sGp / this parity's validity mask no longer needed): hand it back
            tcg_fence_before();
            named_bar_sync(2, TC2_EPI_WARPS * 32);  // every epilogue warp has read sGp before the next tile rewrites it
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
    }

    tcg_fence_before();
    __syncthreads();
    for (int e = tid; e < TI * EM; e += TC2_THREADS) {
        const int il = e / EM, c = e % EM;
        if (i0 + il < L) p.M[size_t(start + i0 + il) * EM + c] = sMsum[e];
    }
    if (warp == 0) {
        tcg_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ K_e3
// f'[r] = W4 SiLU(W3 [f_r, m_r] + b3) + b4 + f_r      (node_mlp + residual, my_egnn_nocoords.py:71-72)
constexpr int NM_ROWS = 16;
// (input and hidden tiles are stored transposed, [k][16 residues], and read with LDS.128)
__global__ void __launch_bounds__(256) embed_node_mlp_kernel(const float* __restrict__ feats, const float* __restrict__ M, int n_res,
                                                             const float* __restrict__ w3t /*[384][256]*/, const float* __restrict__ b3,
                                                             const float* __restrict__ w4t /*[256][128]*/, const float* __restrict__ b4,
                                                             float* __restrict__ feats_out) {
    __shared__ __align__(16) float sInT[EW + EM][NM_ROWS];
    __shared__ __align__(16) float sN1T[EM][NM_ROWS];
    const int r0 = blockIdx.x * NM_ROWS, tid = threadIdx.x;
    for (int e = tid; e < NM_ROWS * (EW + EM); e += 256) {
        const int r = e / (EW + EM), k = e % (EW + EM);
        float v = 0.f;
        if (r0 + r < n_res) v = (k < EW) ? feats[size_t(r0 + r) * EW + k] : M[size_t(r0 + r) * EM + (k - EW)];
        sInT[k][r] = v;
    }
    __syncthreads();
    {
        const int c = tid;  // 256 hidden units
        float acc[NM_ROWS];
        const float b = b3[c];
#pragma unroll
        for (int r = 0; r < NM_ROWS; ++r) acc[r] = b;
#pragma unroll 4
        for (int k = 0; k < EW + EM; ++k) {
            const float w = w3t[size_t(k) * EM + c];
            const float4* f4 = reinterpret_cast<const float4*>(&sInT[k][0]);
#pragma unroll
            for (int r4 = 0; r4 < NM_ROWS / 4; ++r4) {
                const float4 f = f4[r4];
                acc[4 * r4 + 0] = fmaf(f.x, w, acc[4 * r4 + 0]);
                acc[4 * r4 + 1] = fmaf(f.y, w, acc[4 * r4 + 1]);
                acc[4 * r4 + 2] = fmaf(f.z, w, acc[4 * r4 + 2]);
                acc[4 * r4 + 3] = fmaf(f.w, w, acc[4 * r4 + 3]);
            }
        }
#pragma unroll
        for (int r = 0; r < NM_ROWS; ++r) sN1T[c][r] = silu(acc[r]);
    }
    __syncthreads();
    {
        const int c = tid & (EW - 1), half = tid >> 7;  // 128 outputs x 2 halves of 8 residues
        float acc[NM_ROWS / 2];
        const float b = b4[c];
#pragma unroll
        for (int r = 0; r < NM_ROWS / 2; ++r) acc[r] = b;
#pragma unroll 4
        for (int k = 0; k < EM; ++k) {
            const float w = w4t[size_t(k) * EW + c];
            const float4* f4 = reinterpret_cast<const float4*>(&sN1T[k][half * (NM_ROWS / 2)]);
#pragma unroll
            for (int r4 = 0; r4 < NM_ROWS / 8; ++r4) {
                const float4 f = f4[r4];
                acc[4 * r4 + 0] = fmaf(f.x, w, acc[4 * r4 + 0]);
                acc[4 * r4 + 1] = fmaf(f.y, w, acc[4 * r4 + 1]);
                acc[4 * r4 + 2] = fmaf(f.z, w, acc[4 * r4 + 2]);
                acc[4 * r4 + 3] = fmaf(f.w, w, acc[4 * r4 + 3]);
            }
        }
#pragma unroll
        for (int r = 0; r < NM_ROWS / 2; ++r) {
            const int rr = half * (NM_ROWS / 2) + r;
            if (r0 + rr < n_res) feats_out[size_t(r0 + rr) * EW + c] = acc[r] + sInT[c][rr];
        }
    }
}

// ------------------------------------------------------------------------------------------------ K_e4
// embed = out_feats.mean(dim=1)   (nndef_fold_egnn_embed.py:61); one CTA per structure, 4 row groups x 128 columns
__global__ void __launch_bounds__(512) embed_mean_kernel(const float* __restrict__ feats, const int* __restrict__ s_start,
                                                         const int* __restrict__ s_len, float* __restrict__ out) {
    __shared__ float part[4][EW];
    const int s = blockIdx.x, c = threadIdx.x & (EW - 1), g = threadIdx.x >> 7;
    const int start = s_start[s], L = s_len[s];
    float acc = 0.f;
    for (int r = g; r < L; r += 4) acc += feats[size_t(start + r) * EW + c];
    part[g][c] = acc;
    __syncthreads();
    if (g == 0) out[size_t(s) * EW + c] = ((part[0][c] + part[1][c]) + (part[2][c] + part[3][c])) / float(L);
}

// ------------------------------------------------------------------------------------------------ host side
struct LayerDev {
    float* w1abt = nullptr;  // [128][1056]  W1[:, :128]^T | W1[:, 128:256]^T, hidden padded to 528
    float* b1ab = nullptr;   // [1056]       b1 | 0
    float* wd = nullptr;     // [528]
    float* w2t = nullptr;    // [528][256]
    uint8_t* w2img = nullptr; // tensor-core path: [17][hi|lo] bf16 operand images of W2 (32 KB per chunk)
    float* b2 = nullptr;     // [256]
    float* wg = nullptr;     // [256]
    float bg = 0.f;
    float* w3t = nullptr;    // [384][256]
    float* b3 = nullptr;     // [256]
    float* w4t = nullptr;    // [256][128]
    float* b4 = nullptr;     // [128]
};

}  // namespace
}  // namespace fcs

using namespace fcs;

struct fcs_embedder {
    int device = 0;
    int sm_count = 0;
    int n_layers = 0;
    int max_len = 0;
    LayerDev layers[4];
    float* pe = nullptr;
    cudaStream_t stream = nullptr;
    static constexpr int MAX_EDGE_EVENTS = 64;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, edge_ev[2 * MAX_EDGE_EVENTS] = {};
    // workspace (grown on demand)
    int64_t res_cap = 0;
    int struct_cap = 0;
    int64_t item_cap = 0;
    float *coords = nullptr, *feats_a = nullptr, *feats_b = nullptr, *P = nullptr, *Q = nullptr, *M = nullptr, *out = nullptr;
    int *s_start = nullptr, *s_len = nullptr;
    int2* items = nullptr;
    fcs_embed_timing timing = {};
    int mode = FCS_EMBED_MODE_TC3;  // which edge kernel runs (fcs_embed_set_mode; FCS_EMBED_MODE env var at create time)
};

namespace {

int efail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    fcs::set_last_error(buf);
    return code;
}
#define EMB_CUDA(call)                                                                                           \
    do {                                                                                                         \
        cudaError_t e__ = (call);                                                                                \
        if (e__ != cudaSuccess) {                                                                                \
            (void)cudaGetLastError();                                                                            \
            return efail(e__ == cudaErrorMemoryAllocation ? FCS_ERR_NOMEM : FCS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                         cudaGetErrorString(e__), __FILE__, __LINE__);                                           \
        }                                                                                                        \
    } while (0)

struct DevGuard {
    int prev = -1;
    bool ok;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

constexpr int64_t CHUNK_RESIDUES = 1 << 20;  // residues per pass: bounds the workspace at ~6.4 KB x 2^20 = 6.7 GB

int upload(float** dst, const std::vector<float>& host) {
    EMB_CUDA(cudaMalloc(dst, host.size() * sizeof(float)));
    EMB_CUDA(cudaMemcpy(*dst, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    return FCS_OK;
}

int upload_layer(LayerDev& d, const fcs_egnn_weights& w) {
    const int IN1 = 2 * EW + 1;  // 257
    std::vector<float> w1abt(size_t(EW) * 2 * EHP, 0.f), b1ab(2 * EHP, 0.f), wd(EHP, 0.f), w2t(size_t(EHP) * EM, 0.f);
    for (int c = 0; c < EH; ++c) {
        for (int k = 0; k < EW; ++k) {
            w1abt[size_t(k) * 2 * EHP + c] = w.edge_w1[size_t(c) * IN1 + k];
            w1abt[size_t(k) * 2 * EHP + EHP + c] = w.edge_w1[size_t(c) * IN1 + EW + k];
        }
        b1ab[c] = w.edge_b1[c];
        wd[c] = w.edge_w1[size_t(c) * IN1 + 2 * EW];
        for (int m = 0; m < EM; ++m) w2t[size_t(c) * EM + m] = w.edge_w2[size_t(m) * EH + c];
    }
    std::vector<float> b2(w.edge_b2, w.edge_b2 + EM), wg(w.gate_w, w.gate_w + EM);
    std::vector<float> w3t(size_t(EW + EM) * EM), b3(w.node_b1, w.node_b1 + EM), w4t(size_t(EM) * EW), b4(w.node_b2, w.node_b2 + EW);
    for (int c = 0; c < EM; ++c)
        for (int k = 0; k < EW + EM; ++k) w3t[size_t(k) * EM + c] = w.node_w1[size_t(c) * (EW + EM) + k];
    for (int c = 0; c < EW; ++c)
        for (int k = 0; k < EM; ++k) w4t[size_t(k) * EW + c] = w.node_w2[size_t(c) * EM + k];
    d.bg = w.gate_b[0];
    // tensor-core path: W2 [256 out][514 in] as bf16 hi + lo operand images, one 32 KB block per 32-wide chunk of the
    // hidden layer, each part in the canonical K-major no-swizzle core-matrix layout (8 rows x 16 B, LBO 128 B, SBO 512 B)
    std::vector<uint8_t> img(size_t(NCH_T) * B_STAGE, 0);
    for (int c = 0; c < NCH_T; ++c)
        for (int n = 0; n < EM; ++n)
            for (int k = 0; k < KT; ++k) {
                const int kg = c * KT + k;
                const float v = kg < EH ? w.edge_w2[size_t(n) * EH + kg] : 0.f;
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
                const size_t off = size_t(c) * B_STAGE + size_t(n >> 3) * 512 + size_t(k >> 3) * 128 + size_t(n & 7) * 16 + size_t(k & 7) * 2;
                memcpy(&img[off], &hi, 2);
                memcpy(&img[off + B_PART], &lo, 2);
            }
    EMB_CUDA(cudaMalloc(&d.w2img, img.size()));
    EMB_CUDA(cudaMemcpy(d.w2img, img.data(), img.size(), cudaMemcpyHostToDevice));
    int rc;
    if ((rc = upload(&d.w1abt, w1abt)) || (rc = upload(&d.b1ab, b1ab)) || (rc = upload(&d.wd, wd)) || (rc = upload(&d.w2t, w2t)) ||
        (rc = upload(&d.b2, b2)) || (rc = upload(&d.wg, wg)) || (rc = upload(&d.w3t, w3t)) || (rc = upload(&d.b3, b3)) ||
        (rc = upload(&d.w4t, w4t)) || (rc = upload(&d.b4, b4)))
        return rc;
    return FCS_OK;
}

void free_workspace(fcs_embedder* e) {
    cudaFree(e->coords); cudaFree(e->feats_a); cudaFree(e->feats_b); cudaFree(e->P); cudaFree(e->Q); cudaFree(e->M);
    cudaFree(e->out); cudaFree(e->s_start); cudaFree(e->s_len); cudaFree(e->items);
    e->coords = e->feats_a = e->feats_b = e->P = e->Q = e->M = e->out = nullptr;
    e->s_start = e->s_len = nullptr;
    e->items = nullptr;
    e->res_cap = 0; e->struct_cap = 0; e->item_cap = 0;
    (void)cudaGetLastError();
}

int ensure_workspace(fcs_embedder* e, int64_t n_res, int n_structs, int64_t n_items) {
    if (n_res <= e->res_cap && n_structs <= e->struct_cap && n_items <= e->item_cap) return FCS_OK;
    const int64_t rc_ = std::max(n_res, e->res_cap);
    const int sc_ = std::max(n_structs, e->struct_cap);
    const int64_t ic_ = std::max(n_items, e->item_cap);
    free_workspace(e);
    EMB_CUDA(cudaMalloc(&e->coords, size_t(rc_) * 3 * 4));
    EMB_CUDA(cudaMalloc(&e->feats_a, size_t(rc_) * EW * 4));
    EMB_CUDA(cudaMalloc(&e->feats_b, size_t(rc_) * EW * 4));
    EMB_CUDA(cudaMalloc(&e->P, size_t(rc_) * EHP * 4));
    EMB_CUDA(cudaMalloc(&e->Q, size_t(rc_) * EHP * 4));
    EMB_CUDA(cudaMalloc(&e->M, size_t(rc_) * EM * 4));
    EMB_CUDA(cudaMalloc(&e->out, size_t(sc_) * EW * 4));
    EMB_CUDA(cudaMalloc(&e->s_start, size_t(sc_) * 4));
    EMB_CUDA(cudaMalloc(&e->s_len, size_t(sc_) * 4));
    EMB_CUDA(cudaMalloc(&e->items, size_t(ic_) * sizeof(int2)));
    e->res_cap = rc_; e->struct_cap = sc_; e->item_cap = ic_;
    return FCS_OK;
}

// One pass over structures [s0, s1), enqueued on e->stream and synchronised before returning.  The [s1-s0,128] block
// of embeddings goes to `out_dev`, or to the embedder's own buffer e->out when `out_dev` is null and want_out is set.
// If dbg_layer >= 0 the pass stops after that layer (*dbg_feats = that layer's node features, e->M = its messages).
int run_pass(fcs_embedder* e, const float* coords, const int64_t* offsets, int s0, int s1, float* out_dev, bool want_out,
             int* edge_events, int dbg_layer = -1, const float** dbg_feats = nullptr) {
    const int n = s1 - s0;
    const int64_t r_base = offsets[s0], n_res = offsets[s1] - r_base;
    std::vector<int> start(n), len(n), order(n);
    int64_t n_items = 0;
    for (int s = 0; s < n; ++s) {
        start[s] = int(offsets[s0 + s] - r_base);
        len[s] = int(offsets[s0 + s + 1] - offsets[s0 + s]);
        n_items += (len[s] + TI - 1) / TI;
    }
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] > len[b]; });  // longest CTAs first
    std::vector<int2> items;
    items.reserve(size_t(n_items));
    for (int s : order)
        for (int i0 = 0; i0 < len[s]; i0 += TI) items.push_back(make_int2(s, i0));
    int rc = ensure_workspace(e, n_res, n, n_items);
    if (rc != FCS_OK) return rc;
    float* dst = out_dev ? out_dev : (want_out ? e->out : nullptr);
    cudaStream_t st = e->stream;
    // pageable sources: cudaMemcpyAsync returns once they are staged; the stream is synchronised below anyway
    EMB_CUDA(cudaMemcpyAsync(e->coords, coords + r_base * 3, size_t(n_res) * 12, cudaMemcpyHostToDevice, st));
    EMB_CUDA(cudaMemcpyAsync(e->s_start, start.data(), size_t(n) * 4, cudaMemcpyHostToDevice, st));
    EMB_CUDA(cudaMemcpyAsync(e->s_len, len.data(), size_t(n) * 4, cudaMemcpyHostToDevice, st));
    EMB_CUDA(cudaMemcpyAsync(e->items, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice, st));

    embed_init_feats_kernel<<<n, 256, 0, st>>>(e->pe, e->s_start, e->s_len, e->feats_a);
    EMB_CUDA(cudaGetLastError());
    ++e->timing.last_launches;
    float* fin = e->feats_a;
    float* fout = e->feats_b;
    const int row_blocks = int((n_res + NP_ROWS - 1) / NP_ROWS);
    for (int l = 0; l < e->n_layers; ++l) {
        const LayerDev& w = e->layers[l];
        embed_node_proj_kernel<<<row_blocks, 256, 0, st>>>(fin, int(n_res), w.w1abt, w.b1ab, e->P, e->Q);
        EMB_CUDA(cudaGetLastError());
        EdgeParams ep;
        ep.coords = e->coords; ep.s_start = e->s_start; ep.s_len = e->s_len; ep.items = e->items;
        ep.P = e->P; ep.Q = e->Q; ep.wd = w.wd; ep.w2t = w.w2t; ep.b2 = w.b2; ep.wg = w.wg; ep.bg = w.bg; ep.M = e->M;
        ep.w2img = w.w2img;
        const bool timed = *edge_events < fcs_embedder::MAX_EDGE_EVENTS;
        if (timed) EMB_CUDA(cudaEventRecord(e->edge_ev[2 * *edge_events], st));
        if (e->mode == FCS_EMBED_MODE_TC2)
            embed_edge_tc2_kernel<8><<<int(n_items), tc2_threads(8), EDGE_TC2_SMEM, st>>>(ep);
        else if (e->mode == FCS_EMBED_MODE_TC3)
            embed_edge_tc2_kernel<16><<<int(n_items), tc2_threads(16), EDGE_TC2_SMEM, st>>>(ep);
        else if (e->mode == FCS_EMBED_MODE_TC)
            embed_edge_tc_kernel<<<int(n_items), TC_THREADS_ALL, EDGE_TC_SMEM, st>>>(ep);
        else
            embed_edge_kernel<<<int(n_items), EDGE_THREADS, EDGE_SMEM, st>>>(ep);
        EMB_CUDA(cudaGetLastError());
        if (timed) {
            EMB_CUDA(cudaEventRecord(e->edge_ev[2 * *edge_events + 1], st));
            ++*edge_events;
        }
        embed_node_mlp_kernel<<<(int(n_res) + NM_ROWS - 1) / NM_ROWS, 256, 0, st>>>(fin, e->M, int(n_res), w.w3t, w.b3, w.w4t, w.b4, fout);
        EMB_CUDA(cudaGetLastError());
        e->timing.last_launches += 3;
        std::swap(fin, fout);
        if (l == dbg_layer) {
            *dbg_feats = fin;
            EMB_CUDA(cudaStreamSynchronize(st));
            return FCS_OK;
        }
    }
    if (dst) {
        embed_mean_kernel<<<n, 512, 0, st>>>(fin, e->s_start, e->s_len, dst);
        EMB_CUDA(cudaGetLastError());
        ++e->timing.last_launches;
    }
    // the next pass reuses the workspace, and the pageable staging vectors die with this frame
    EMB_CUDA(cudaStreamSynchronize(st));
    return FCS_OK;
}

int validate(const fcs_embedder* e, const float* coords, const int64_t* offsets, int n, const void* out) {
    if (!e) return efail(FCS_ERR_INVALID, "fcs_embed: null embedder");
    if (n < 0 || (n > 0 && (!coords || !offsets || !out))) return efail(FCS_ERR_INVALID, "fcs_embed: null argument");
    for (int s = 0; s < n; ++s) {
        const int64_t L = offsets[s + 1] - offsets[s];
        if (L < 1 || L > e->max_len)
            return efail(FCS_ERR_INVALID, "fcs_embed: structure %d has %lld residues; 1..%d supported (positional table, "
                         "nndef_fold_egnn_embed.py:13)", s, (long long)L, e->max_len);
    }
    if (n > 0 && offsets[0] < 0) return efail(FCS_ERR_INVALID, "fcs_embed: negative offset");
    return FCS_OK;
}

// out_dev == nullptr: results go to out_host through e->out
int embed_impl(fcs_embedder* e, const float* coords, const int64_t* offsets, int n, float* out_host, float* out_dev) {
    int rc = validate(e, coords, offsets, n, out_host ? (const void*)out_host : (const void*)out_dev);
    if (rc != FCS_OK) return rc;
    DevGuard guard(e->device);
    if (!guard.ok) return efail(FCS_ERR_CUDA, "fcs_embed: cudaSetDevice(%d) failed", e->device);
    e->timing = fcs_embed_timing{};
    e->timing.last_structures = n;
    if (n == 0) return FCS_OK;
    e->timing.last_residues = offsets[n] - offsets[0];
    for (int s = 0; s < n; ++s) {
        const int64_t L = offsets[s + 1] - offsets[s];
        e->timing.last_pairs += L * L;
    }
    EMB_CUDA(cudaEventRecord(e->ev0, e->stream));
    int edge_events = 0;
    for (int s0 = 0; s0 < n;) {
        int s1 = s0 + 1;
        while (s1 < n && offsets[s1 + 1] - offsets[s0] <= CHUNK_RESIDUES) ++s1;
        rc = run_pass(e, coords, offsets, s0, s1, out_dev ? out_dev + size_t(s0) * EW : nullptr, true, &edge_events);
        if (rc != FCS_OK) return rc;
        if (!out_dev) {
            EMB_CUDA(cudaMemcpyAsync(out_host + size_t(s0) * EW, e->out, size_t(s1 - s0) * EW * 4, cudaMemcpyDeviceToHost, e->stream));
            EMB_CUDA(cudaStreamSynchronize(e->stream));
        }
        s0 = s1;
    }
    EMB_CUDA(cudaEventRecord(e->ev1, e->stream));
    EMB_CUDA(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->timing.last_ms = ms;
    for (int i = 0; i < edge_events; ++i)
        if (cudaEventElapsedTime(&ms, e->edge_ev[2 * i], e->edge_ev[2 * i + 1]) == cudaSuccess) e->timing.last_edge_ms += ms;
    return FCS_OK;
}

}  // namespace

extern "C" int fcs_embedder_create(int device, const fcs_egnn_weights* layers, int n_layers, const float* pe, int max_len,
                                   fcs_embedder** out) {
    if (!out) return efail(FCS_ERR_INVALID, "fcs_embedder_create: out is null");
    *out = nullptr;
    if (!layers || !pe || n_layers < 1 || n_layers > 4 || max_len < 1)
        return efail(FCS_ERR_INVALID, "fcs_embedder_create: need 1..4 layers, a positional table and max_len >= 1");
    for (int l = 0; l < n_layers; ++l) {
        const fcs_egnn_weights& w = layers[l];
        if (!w.edge_w1 || !w.edge_b1 || !w.edge_w2 || !w.edge_b2 || !w.gate_w || !w.gate_b || !w.node_w1 || !w.node_b1 ||
            !w.node_w2 || !w.node_b2)
            return efail(FCS_ERR_INVALID, "fcs_embedder_create: layer %d has a null weight pointer", l);
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        (void)cudaGetLastError();
        return efail(FCS_ERR_CUDA, "fcs_embedder_create: no usable CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= count) return efail(FCS_ERR_INVALID, "fcs_embedder_create: device %d out of range", device);
    DevGuard guard(device);
    if (!guard.ok) return efail(FCS_ERR_CUDA, "fcs_embedder_create: cudaSetDevice(%d) failed", device);
    fcs_embedder* e = new (std::nothrow) fcs_embedder();
    if (!e) return efail(FCS_ERR_NOMEM, "fcs_embedder_create: out of host memory");
    e->device = device;
    e->n_layers = n_layers;
    e->max_len = max_len < FCS_EMBED_MAX_LEN ? max_len : FCS_EMBED_MAX_LEN;
    auto run = [&]() -> int {
        cudaDeviceProp prop;
        EMB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) return efail(FCS_ERR_UNSUPPORTED, "fcs_embedder_create: built for sm_100a, device is sm_%d%d", prop.major, prop.minor);
        e->sm_count = prop.multiProcessorCount;
        EMB_CUDA(cudaFuncSetAttribute(embed_edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_SMEM));
        EMB_CUDA(cudaFuncSetAttribute(embed_edge_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_TC_SMEM));
        EMB_CUDA(cudaFuncSetAttribute(embed_edge_tc2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_TC2_SMEM));
        EMB_CUDA(cudaFuncSetAttribute(embed_edge_tc2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_TC2_SMEM));
        if (const char* m = getenv("FCS_EMBED_MODE")) {
            const int v = atoi(m);
            if (v >= FCS_EMBED_MODE_FP32 && v <= FCS_EMBED_MODE_TC3) e->mode = v;
        }
        EMB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        EMB_CUDA(cudaEventCreate(&e->ev0));
        EMB_CUDA(cudaEventCreate(&e->ev1));
        for (cudaEvent_t& ev : e->edge_ev) EMB_CUDA(cudaEventCreate(&ev));
        std::vector<float> pe_h(pe, pe + size_t(e->max_len) * EW);
        int rc = upload(&e->pe, pe_h);
        if (rc != FCS_OK) return rc;
        for (int l = 0; l < n_layers; ++l)
            if ((rc = upload_layer(e->layers[l], layers[l])) != FCS_OK) return rc;
        return FCS_OK;
    };
    const int rc = run();
    if (rc != FCS_OK) {
        fcs_embedder_destroy(e);
        return rc;
    }
    *out = e;
    return FCS_OK;
}

extern "C" int fcs_embedder_destroy(fcs_embedder* e) {
    if (!e) return FCS_OK;
    DevGuard guard(e->device);
    free_workspace(e);
    for (LayerDev& d : e->layers) {
        cudaFree(d.w1abt); cudaFree(d.b1ab); cudaFree(d.wd); cudaFree(d.w2t); cudaFree(d.b2); cudaFree(d.wg);
        cudaFree(d.w3t); cudaFree(d.b3); cudaFree(d.w4t); cudaFree(d.b4); cudaFree(d.w2img);
    }
    cudaFree(e->pe);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    for (cudaEvent_t ev : e->edge_ev)
        if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    (void)cudaGetLastError();
    delete e;
    return FCS_OK;
}

extern "C" int fcs_embed(fcs_embedder* e, const float* coords, const int64_t* offsets, int n, float* out_host) {
    return embed_impl(e, coords, offsets, n, out_host, nullptr);
}

extern "C" int fcs_embed_to_device(fcs_embedder* e, const float* coords, const int64_t* offsets, int n, float* out_dev) {
    if (n > 0 && !out_dev) return efail(FCS_ERR_INVALID, "fcs_embed_to_device: out_dev is null");
    return embed_impl(e, coords, offsets, n, nullptr, out_dev);
}

extern "C" int fcs_embed_set_mode(fcs_embedder* e, int mode) {
    if (!e) return efail(FCS_ERR_INVALID, "fcs_embed_set_mode: null embedder");
    if (mode < FCS_EMBED_MODE_FP32 || mode > FCS_EMBED_MODE_TC3) return efail(FCS_ERR_INVALID, "fcs_embed_set_mode: unknown mode %d", mode);
    e->mode = mode;
    return FCS_OK;
}

extern "C" int fcs_embed_get_timing(const fcs_embedder* e, fcs_embed_timing* out) {
    if (!e || !out) return efail(FCS_ERR_INVALID, "fcs_embed_get_timing: null argument");
    *out = e->timing;
    return FCS_OK;
}

extern "C" int fcs_embed_debug_layer(fcs_embedder* e, const float* coords, int length, int layer, float* out_feats, float* out_messages) {
    if (!e || !coords) return efail(FCS_ERR_INVALID, "fcs_embed_debug_layer: null argument");
    if (layer < 0 || layer >= e->n_layers) return efail(FCS_ERR_INVALID, "fcs_embed_debug_layer: layer %d out of range", layer);
    const int64_t offsets[2] = {0, length};
    int rc = validate(e, coords, offsets, 1, coords);
    if (rc != FCS_OK) return rc;
    DevGuard guard(e->device);
    if (!guard.ok) return efail(FCS_ERR_CUDA, "fcs_embed_debug_layer: cudaSetDevice(%d) failed", e->device);
    e->timing = fcs_embed_timing{};
    int edge_events = 0;
    const float* feats = nullptr;
    rc = run_pass(e, coords, offsets, 0, 1, nullptr, false, &edge_events, layer, &feats);
    if (rc != FCS_OK) return rc;
    EMB_CUDA(cudaStreamSynchronize(e->stream));
    if (out_feats) EMB_CUDA(cudaMemcpy(out_feats, feats, size_t(length) * EW * 4, cudaMemcpyDeviceToHost));
    if (out_messages) EMB_CUDA(cudaMemcpy(out_messages, e->M, size_t(length) * EM * 4, cudaMemcpyDeviceToHost));
    return FCS_OK;
}
