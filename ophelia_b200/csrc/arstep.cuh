// Autoregressive frame step (synthesize.py:150-230) computed incrementally.
//
// The reference re-runs AudioEnc + Attention + AudioDec over all max_T frames for every generated frame and keeps only
// row j of the result.  AudioEnc and AudioDec are stacks of CAUSAL convolutions (networks.py:214-284, 360-435), so row j of
// every layer depends only on rows j, j - rate, j - 2 rate of the layer below, all of which were already computed in the
// earlier steps.  Every layer therefore keeps its output history [B][T][C] in HBM and a frame step computes ONE row per
// layer and item: B x (k Cin) x Cout multiply-adds instead of B x T x (k Cin) x Cout.  With B <= 16 rows this is a
// weight-streaming (L2 / HBM bound) job, not a tensor-core one: the fp32 kernel [k][Cin][Cout] is read once per step
// with coalesced loads along Cout, every weight is used for all B items from a register, and the reduction over k Cin is
// cut into slices spread over the SMs.  Two launches per layer:
//   ar_gemv_kernel   partial[s][b][o] = sum over slice s of x[b][tap row][c] * W[tap][c][o]
//   ar_tail_kernel   z = bias + sum_s partial (fixed order: deterministic), LayerNorm / ReLU / sigmoid, or the highway
//                    mix with the layer's own input row; writes row j of the layer's history
// This caching is exact for AudioEnc only (its input row t is final once frame t - 1 exists); see the window kernels
// below for Attention / AudioDec.  ar_advance_kernel keeps the frame counter on the device, so that ONE captured CUDA
// graph serves every frame.
// All arithmetic is fp32 FMA (the tcgen05 path's split-bf16 products are an emulation of exactly this).
#pragma once
#include "rowwise.cuh"
#include <cooperative_groups.h>

namespace oph {

constexpr int AR_MAXB = 16;          // items per frame step
constexpr int AR_COLS = 64;          // output channels per CTA
constexpr int AR_SUBS = 4;           // k sub-slices per CTA (256 threads)
constexpr int AR_KSLICE_MAX = 256;   // reduction length staged per CTA

// grid (ceil(O / 64), KS), block 256.  x: history of the layer input [B][T][ldx] (item stride x_item floats); the row
// read for tap i is  t_i = j - in_shift - (k - 1 - i) * rate  (causal padding, modules.py:123-127), zero when t_i < 0.
__global__ void ar_gemv_kernel(const float* __restrict__ W, const float* __restrict__ x, long long x_item, long long ldx,
                               int Cin, int O, int k, int rate, int in_shift, const int* __restrict__ frame, int B,
                               float* __restrict__ partial, int kslice) {
    pdl_grid_sync();
    __shared__ float xs[AR_MAXB][AR_KSLICE_MAX];
    __shared__ float red[AR_SUBS][AR_MAXB][AR_COLS];
    const int j = *frame;
    const int K = k * Cin;
    const int k0 = blockIdx.y * kslice;
    const int k1 = min(K, k0 + kslice);
    const int len = k1 - k0;
    for (int idx = threadIdx.x; idx < B * len; idx += blockDim.x) {
        const int b = idx / len, kk = idx - b * len + k0;
        const int tap = kk / Cin, c = kk - tap * Cin;
        const int t = j - in_shift - (k - 1 - tap) * rate;
        xs[b][kk - k0] = t >= 0 ? x[(long long)b * x_item + (long long)t * ldx + c] : 0.f;
    }
    __syncthreads();
    const int col = blockIdx.x * AR_COLS + (threadIdx.x & (AR_COLS - 1));
    const int sub = threadIdx.x / AR_COLS;
    float acc[AR_MAXB];
#pragma unroll
    for (int b = 0; b < AR_MAXB; ++b) acc[b] = 0.f;
    if (col < O) {
        for (int kk = sub; kk < len; kk += AR_SUBS) {
            const float w = __ldg(W + (long long)(k0 + kk) * O + col);
#pragma unroll
            for (int b = 0; b < AR_MAXB; ++b)
                if (b < B) acc[b] = fmaf(w, xs[b][kk], acc[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < AR_MAXB; ++b) red[sub][b][threadIdx.x & (AR_COLS - 1)] = acc[b];
    __syncthreads();
    for (int idx = threadIdx.x; idx < B * AR_COLS; idx += blockDim.x) {
        const int b = idx / AR_COLS, cc = idx - b * AR_COLS;
        const int o = blockIdx.x * AR_COLS + cc;
        if (o < O) {
            float s = red[0][b][cc];
#pragma unroll
            for (int q = 1; q < AR_SUBS; ++q) s += red[q][b][cc];
            partial[((long long)blockIdx.y * B + b) * O + o] = s;
        }
    }
}

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {       // blockDim.x == 256; scratch[9]
    v = warp_sum(v);
    __syncthreads();                                   // scratch may still be read from the previous reduction
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? scratch[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) scratch[8] = t;
    }
    __syncthreads();
    return scratch[8];
}

// LayerNorm of zs[0..C) in shared memory, in place (tf.contrib.layers.layer_norm: biased variance, eps 1e-12; two passes)
__device__ __forceinline__ void block_layer_norm(float* zs, int C, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float* scratch) {
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += zs[c];
    const float mean = block_sum_256(s, scratch) / (float)C;
    float q = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { const float d = zs[c] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum_256(q, scratch) / (float)C + LN_EPS);
    for (int c = threadIdx.x; c < C; c += blockDim.x) zs[c] = (zs[c] - mean) * rstd * gamma[c] + beta[c];
    __syncthreads();
}

// grid B, block 256, dynamic shared memory O floats.  kind 0: conv1d tail (modules.py:137-141; y_sig = sigmoid of the
// normalised logits, networks.py:430-433); kind 1: highway tail (modules.py:194-205), O = 2 C, xres = the layer's own input.
__global__ void ar_tail_kernel(const float* __restrict__ partial, int KS, const float* __restrict__ bias, int O, int kind,
                               const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
                               const float* __restrict__ b2, int act, const float* __restrict__ xres, long long x_item,
                               long long ldx, float* __restrict__ y, long long y_item, long long ldy,
                               float* __restrict__ y_sig, long long s_item, long long lds, const int* __restrict__ frame,
                               int B) {
    pdl_grid_sync();
    extern __shared__ float zs[];
    __shared__ float scratch[9];
    const int b = blockIdx.x, j = *frame;
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        float s = bias ? bias[o] : 0.f;
        for (int q = 0; q < KS; ++q) s += partial[((long long)q * B + b) * O + o];
        zs[o] = s;
    }
    __syncthreads();
    if (kind == 0) {
        if (g1) block_layer_norm(zs, O, g1, b1, scratch);
        float* yr = y + (long long)b * y_item + (long long)j * ldy;
        float* sr = y_sig ? y_sig + (long long)b * s_item + (long long)j * lds : nullptr;
        for (int o = threadIdx.x; o < O; o += blockDim.x) {
            const float u = zs[o];
            yr[o] = act ? fmaxf(u, 0.f) : u;
            if (sr) sr[o] = sigmoidf_(u);
        }
    } else {
        const int C = O / 2;
        if (g1) {
            block_layer_norm(zs, C, g1, b1, scratch);
            block_layer_norm(zs + C, C, g2, b2, scratch);
        }
        const float* xr = xres + (long long)b * x_item + (long long)j * ldx;
        float* yr = y + (long long)b * y_item + (long long)j * ldy;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float g = sigmoidf_(zs[c]);
            yr[c] = g * zs[C + c] + (1.f - g) * xr[c];
        }
    }
}

// Attention is not cacheable: the window mask of the latest prev_max_attentions applies to every time row
// (networks.py:304-313), so R[t < j] changes when the window moves and AudioDec's row j sees those rows through its
// causal reach (84 frames).  Attention and AudioDec therefore run per step with the batch kernels, but only over the W =
// reach + 1 rows [s, s + W), s = max(0, min(j - reach, T - W)); these two kernels move that window in and out of the
// histories (the frame index lives on the device, the window buffers are static: one CUDA graph serves every frame).
__device__ __forceinline__ int ar_window_start(int j, int reach, int T, int W) {
    return max(0, min(j - reach, T - W));
}

// grid (W, B), block 128:  Qw[b][w][:] = Q[b][s + w][:]
__global__ void ar_window_gather_kernel(const float* __restrict__ Q, long long q_item, long long ldq,
                                        float* __restrict__ Qw, long long w_item, long long ldw, int d, int T, int W,
                                        int reach, const int* __restrict__ frame) {
    pdl_grid_sync();
    const int s = ar_window_start(*frame, reach, T, W);
    const int w = blockIdx.x, b = blockIdx.y;
    const float* src = Q + (long long)b * q_item + (long long)(s + w) * ldq;
    float* dst = Qw + (long long)b * w_item + (long long)w * ldw;
    for (int c = threadIdx.x; c < d; c += blockDim.x) dst[c] = src[c];
}

// grid B, block 256:  row r = j - s of the window results becomes frame j:  Y[b][j][:], alignments[b][:][j], the new
// prev_max_attentions[b] and the argmax history [T][B] (synthesize.py:204-209)
__global__ void ar_window_scatter_kernel(const float* __restrict__ Yw, long long yw_item, long long ldyw,
                                         float* __restrict__ Y, long long y_item, long long ldy, int nm,
                                         const float* __restrict__ align_w, float* __restrict__ align_t,
                                         const int* __restrict__ argmax_w, int* __restrict__ prev,
                                         int* __restrict__ history, int B, int N, int T, int W, int reach,
                                         const int* __restrict__ frame) {
    pdl_grid_sync();
    const int j = *frame, b = blockIdx.x;
    const int r = j - ar_window_start(j, reach, T, W);
    const float* yw = Yw + (long long)b * yw_item + (long long)r * ldyw;
    float* yr = Y + (long long)b * y_item + (long long)j * ldy;
    for (int c = threadIdx.x; c < nm; c += blockDim.x) yr[c] = yw[c];
    for (int n = threadIdx.x; n < N; n += blockDim.x)
        align_t[((long long)b * N + n) * T + j] = align_w[((long long)b * N + n) * W + r];
    if (threadIdx.x == 0) {
        const int a = argmax_w[(long long)b * W + r];
        prev[b] = a;
        history[(long long)j * B + b] = a;
    }
}

// ---- the whole AudioEnc frame step in ONE launch ---------------------------------------------------------------
// Row j of every AudioEnc layer depends on that item's own histories only, so the 13 layers of an item can run back to
// back inside one kernel.  A single SM cannot stream the 16 MB of weights fast enough (~120 GB/s from L2), so an item is
// given a thread-block CLUSTER of 8 CTAs: CTA r owns the output channels [r C/8, (r+1) C/8) of every layer (for a highway
// layer the matching slices of H1 and H2), the LayerNorm moments are exchanged through distributed shared memory, and a
// cluster barrier publishes row j of a layer (written to its history in HBM / L2) to the 7 peers before the next layer
// stages it.  26 launches per frame become one (0.388 against 0.413 ms per frame step at 10 sentences on B200).
struct ArEncLayer {
    const float* w; const float* bias; const float* g1; const float* b1; const float* g2; const float* b2;
    const float* x; float* y;
    long long x_item, ldx, y_item, ldy;
    int Cin, C, k, rate, kind, act, in_shift;      // kind 0: conv1d (C = Cout), 1: highway (C channels, 2 C conv outputs)
};
constexpr int AR_ENC_MAX_LAYERS = 16;
constexpr int AR_ENC_CLUSTER = 8;
constexpr int AR_ENC_MAXK = 3072;
struct ArEncArgs { ArEncLayer L[AR_ENC_MAX_LAYERS]; int nlayers; };

constexpr int AR_ENC_THREADS = 1024;

__device__ __forceinline__ void ar_cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(AR_ENC_CLUSTER, 1, 1) __launch_bounds__(AR_ENC_THREADS)
ar_encoder_kernel(const ArEncArgs a, const int* __restrict__ frame) {
    namespace cg = cooperative_groups;
    pdl_grid_sync();
    cg::cluster_group cl = cg::this_cluster();
    __shared__ float xs[AR_ENC_MAXK];
    __shared__ float red[AR_ENC_THREADS];            // [k sub-slice][column]
    __shared__ float zs[128];
    __shared__ float2 part[2][2];                    // [layer parity][H1 / H2]: (mean, sum of squared deviations) of this CTA's columns
    const int r = (int)cl.block_rank(), b = blockIdx.y, tid = threadIdx.x;
    const int j = *frame;
    for (int l = 0; l < a.nlayers; ++l) {
        const ArEncLayer& Ly = a.L[l];
        const int K = Ly.k * Ly.Cin, C = Ly.C, hc = Ly.kind, cpc = C / AR_ENC_CLUSTER;
        const int ncols = hc ? 2 * cpc : cpc, O = hc ? 2 * C : C, nsub = AR_ENC_THREADS / ncols, ph = l & 1;
        for (int idx = tid; idx < K; idx += AR_ENC_THREADS) {
            const int tap = idx / Ly.Cin, c = idx - tap * Ly.Cin;
            const int t = j - Ly.in_shift - (Ly.k - 1 - tap) * Ly.rate;
            // (row j of the layer below was written by the peer CTAs a moment ago: read through L2)
            xs[idx] = t >= 0 ? __ldcg(Ly.x + (long long)b * Ly.x_item + (long long)t * Ly.ldx + c) : 0.f;
        }
        __syncthreads();
        const int lc = tid % ncols, sub = tid / ncols;
        // channel of this thread's column inside its group, and the conv output column it accumulates
        const int ch = r * cpc + (lc < cpc ? lc : lc - cpc);
        const int gcol = (hc && lc >= cpc) ? C + ch : ch;
        {
            // four independent chains, four weight loads in flight per thread (the weights come from L2)
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const float* wp = Ly.w + gcol;
            int kk = sub;
            for (; kk + 3 * nsub < K; kk += 4 * nsub) {
                const float w0 = __ldg(wp + (long long)kk * O), w1 = __ldg(wp + (long long)(kk + nsub) * O);
                const float w2 = __ldg(wp + (long long)(kk + 2 * nsub) * O), w3 = __ldg(wp + (long long)(kk + 3 * nsub) * O);
                a0 = fmaf(w0, xs[kk], a0); a1 = fmaf(w1, xs[kk + nsub], a1);
                a2 = fmaf(w2, xs[kk + 2 * nsub], a2); a3 = fmaf(w3, xs[kk + 3 * nsub], a3);
            }
            for (; kk < K; kk += nsub) a0 = fmaf(__ldg(wp + (long long)kk * O), xs[kk], a0);
            red[sub * ncols + lc] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        if (tid < ncols) {
            float s = Ly.bias ? Ly.bias[gcol] : 0.f;
            for (int q = 0; q < nsub; ++q) s += red[q * ncols + tid];
            zs[tid] = s;
        }
        __syncthreads();
        const int grp = (hc && lc >= cpc) ? 1 : 0;    // which LayerNorm a column of this CTA belongs to
        float u = tid < ncols ? zs[tid] : 0.f;
        if (Ly.g1) {
            // moments of the C channels of a group from the 8 CTAs' (mean, M2) pairs (two passes locally, Chan's
            // combination across the cluster: as accurate as two global passes, one exchange instead of two)
            if (tid < (hc ? 2 : 1)) {
                float m = 0.f;
                for (int c = 0; c < cpc; ++c) m += zs[tid * cpc + c];
                m /= (float)cpc;
                float q = 0.f;
                for (int c = 0; c < cpc; ++c) { const float d = zs[tid * cpc + c] - m; q += d * d; }
                part[ph][tid] = make_float2(m, q);
            }
            ar_cluster_barrier();
            float2 pr[AR_ENC_CLUSTER];
            float mean = 0.f;
#pragma unroll
            for (int rr = 0; rr < AR_ENC_CLUSTER; ++rr) { pr[rr] = cl.map_shared_rank(&part[ph][0], rr)[grp]; mean += pr[rr].x; }
            mean *= 1.f / (float)AR_ENC_CLUSTER;
            float m2 = 0.f;
#pragma unroll
            for (int rr = 0; rr < AR_ENC_CLUSTER; ++rr) { const float d = pr[rr].x - mean; m2 += pr[rr].y + (float)cpc * d * d; }
            const float rstd = rsqrtf(m2 / (float)C + LN_EPS);
            if (tid < ncols) {
                const float gam = grp ? Ly.g2[ch] : Ly.g1[ch], bet = grp ? Ly.b2[ch] : Ly.b1[ch];
                u = (u - mean) * rstd * gam + bet;
            }
        }
        float* yr = Ly.y + (long long)b * Ly.y_item + (long long)j * Ly.ldy;
        if (!hc) {
            if (tid < ncols) yr[ch] = Ly.act ? fmaxf(u, 0.f) : u;
        } else {
            __syncthreads();                          // every column (and the moment threads) has read zs
            if (tid < ncols) zs[tid] = u;
            __syncthreads();
            if (tid < cpc) {
                const float g = sigmoidf_(zs[tid]);
                const float xres = xs[(Ly.k - 1) * Ly.Cin + ch];      // tap k - 1 is row j of the layer input
                yr[ch] = g * zs[cpc + tid] + (1.f - g) * xres;
            }
        }
        ar_cluster_barrier();                         // row j of this layer is visible to the cluster; xs / zs / red are free
    }
}

__global__ void ar_advance_kernel(int* frame) {
    pdl_grid_sync();
    *frame += 1;
}

}  // namespace oph
