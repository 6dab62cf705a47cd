// Row-wise (HBM-bound) kernels of the dc_tts path: layer-norm / activation / highway mix (+ their backward),
// softmax with the monotonic window mask and guided-attention loss, embedding, losses, Adam.
// Generic kernels: one warp owns one [C]-row, reductions along channels are warp shuffles, loads are lane-contiguous.
// Hot shapes (C = 256 / 512 / 1024) use the vectorised / streaming variants further down (float4 per lane, rows
// prefetched through per-warp cp.async rings, WPR warps per row).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace oph {

constexpr float LN_EPS = 1e-12f;                 // tf.contrib.layers.layer_norm (modules.py:65)
constexpr float ATT_MASK_VALUE = -4294967295.0f; // -2**32+1 (networks.py:312)

// Programmatic dependent launch: wait for the predecessor kernel of the stream (memory visible afterwards), then
// allow the successor's blocks to become resident early.  A no-op for launches without the PDL attribute.
__device__ __forceinline__ void pdl_grid_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// (fast reciprocal: 2 ulp, far below the 2e-5 noise of the split-bf16 products; the IEEE division costs ~8 instructions
// per element in kernels that are issue-bound as much as memory-bound)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// counter-based dropout mask (recomputed in backward): keep iff u >= rate; kept values scaled by 1/(1-rate)
// (32-bit avalanche hash of (seed, element index): a handful of integer instructions per element)
__device__ __forceinline__ float drop_scale(unsigned long long seed, unsigned long long idx, float rate, float inv_keep) {
    uint32_t h = (uint32_t)idx * 0x9E3779B1u + (uint32_t)seed;
    h ^= ((uint32_t)(idx >> 32) * 0x85EBCA77u) ^ (uint32_t)(seed >> 32);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    return u >= rate ? inv_keep : 0.f;
}
__device__ __forceinline__ unsigned long long eff_seed(unsigned long long seed, const long long* step) {
    return step ? seed + (unsigned long long)(*step) * 0xD1342543DE82EF95ull : seed;
}

__device__ __forceinline__ void split4(const float4& v, uint2& hh, uint2& ll) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    hh.x = *reinterpret_cast<const uint32_t*>(&h0); hh.y = *reinterpret_cast<const uint32_t*>(&h1);
    ll.x = *reinterpret_cast<const uint32_t*>(&l0); ll.y = *reinterpret_cast<const uint32_t*>(&l1);
}
__device__ __forceinline__ void st_split1(unsigned short* hi, unsigned short* lo, long long idx, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi[idx] = *reinterpret_cast<const unsigned short*>(&h);
    lo[idx] = *reinterpret_cast<const unsigned short*>(&l);
}

struct RowStats { float mean, rstd; };

// two-pass moments like tf.nn.moments: mean, then mean of squared deviations (biased)
__device__ __forceinline__ RowStats row_stats(const float* __restrict__ z, int C, int lane) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += z[c];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = z[c] - mean; q += d * d; }
    const float var = warp_sum(q) / (float)C;
    return {mean, rsqrtf(var + LN_EPS)};
}

// ------------------------------------------------------------------------------------------------
// conv1d tail (modules.py:137-141): y = dropout(act(LN(z)));  optional y_sig = sigmoid(LN(z)) (networks.py:430-433)
__global__ void ln_act_fwd_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ y, long long ldy,
                                  float* __restrict__ y_sig, long long ldys, unsigned short* __restrict__ y_hi,
                                  unsigned short* __restrict__ y_lo, long long ldp, float* __restrict__ stats,
                                  int rows, int C, int act, int norm, float drop_p, unsigned long long seed,
                                  const long long* step) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned long long sd = eff_seed(seed, step);
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const float* zr = z + row * ldz;
        RowStats st = {0.f, 1.f};
        if (norm) st = row_stats(zr, C, lane);
        if (stats && lane == 0) { stats[row * 2] = st.mean; stats[row * 2 + 1] = st.rstd; }
        for (int c = lane; c < C; c += 32) {
            float u = zr[c];
            if (norm) u = (u - st.mean) * st.rstd * gamma[c] + beta[c];
            if (y_sig) y_sig[row * ldys + c] = sigmoidf_(u);
            float a = act == 1 ? fmaxf(u, 0.f) : u;
            if (drop_p > 0.f) a *= drop_scale(sd, (unsigned long long)row * C + c, drop_p, inv_keep);
            y[row * ldy + c] = a;
            if (y_hi) st_split1(y_hi, y_lo, row * ldp + c, a);
        }
    }
}

// highway tail (modules.py:194-205): z = [H1 | H2]; g = sigmoid(LN1(H1)); h = LN2(H2); y = dropout(g*h + (1-g)*x)
// Per-speaker channel gates (modules.learn_channel_contributions, modules.py:78-88): gate[b][c] = sigmoid(table[code_b][c])
// with the zero-padded embedding (code 0 reads as zeros -> 0.5); L = rows per batch item.  table == NULL: no gates.
struct LccGate { const float* table; const int* codes; int L; float* scratch; };
__device__ __forceinline__ float lcc_gate(const LccGate& g, long long row, int c, int C) {
    const int code = g.codes[row / g.L];
    return code == 0 ? 0.5f : sigmoidf_(__ldg(g.table + (long long)code * C + c));
}

__global__ void hc_post_fwd_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ x, long long ldx,
                                   const float* __restrict__ g1, const float* __restrict__ b1,
                                   const float* __restrict__ g2, const float* __restrict__ b2,
                                   float* __restrict__ y, long long ldy, float* __restrict__ stats,
                                   int rows, int C, int norm, float drop_p, unsigned long long seed, const long long* step,
                                   LccGate lcc, unsigned short* __restrict__ y_hi, unsigned short* __restrict__ y_lo, long long ldp) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned long long sd = eff_seed(seed, step);
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const float* z1 = z + row * ldz;
        const float* z2 = z1 + C;
        const float* xr = x + row * ldx;
        RowStats s1 = {0.f, 1.f}, s2 = {0.f, 1.f};
        if (norm) { s1 = row_stats(z1, C, lane); s2 = row_stats(z2, C, lane); }
        if (stats && lane == 0) {
            stats[row * 4] = s1.mean; stats[row * 4 + 1] = s1.rstd; stats[row * 4 + 2] = s2.mean; stats[row * 4 + 3] = s2.rstd;
        }
        for (int c = lane; c < C; c += 32) {
            float u1 = z1[c], u2 = z2[c];
            if (norm) {
                u1 = (u1 - s1.mean) * s1.rstd * g1[c] + b1[c];
                u2 = (u2 - s2.mean) * s2.rstd * g2[c] + b2[c];
            }
            const float g = sigmoidf_(u1);
            if (lcc.table) u2 *= lcc_gate(lcc, row, c, C);              // LCC on the transformation connection only (modules.py:200-201)
            float o = g * u2 + (1.f - g) * xr[c];
            if (drop_p > 0.f) o *= drop_scale(sd, (unsigned long long)row * C + c, drop_p, inv_keep);
            y[row * ldy + c] = o;
            if (y_hi) st_split1(y_hi, y_lo, row * ldp + c, o);
        }
    }
}

// backward of ln_act: dz (pre-LN conv output gradient); column sums dgamma/dbeta/dbias via shared atomics.
// Two passes over the row, nothing staged in dz: the output may be fp32 or split-bf16 planes (dz_hi != NULL).
__global__ void ln_act_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ z, long long ldz,
                                  const float* __restrict__ stats, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ dz, long long lddz,
                                  unsigned short* __restrict__ dz_hi, unsigned short* __restrict__ dz_lo, long long ldp,
                                  float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                                  int rows, int C, int act, int norm, float drop_p, unsigned long long seed,
                                  const long long* step) {
    pdl_grid_sync();
    extern __shared__ float sacc[];            // [3][C]: dgamma, dbeta, dbias
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float invC = 1.f / (float)C;
    const unsigned long long sd = eff_seed(seed, step);
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const float* zr = z + row * ldz;
        const float* dyr = dy + row * lddy;
        float mean = 0.f, rstd = 1.f;
        if (norm) { mean = stats[row * 2]; rstd = stats[row * 2 + 1]; }
        auto du_at = [&](int c, float& xh) {       // gradient w.r.t. the LN output u (after dropout / ReLU backward)
            xh = norm ? (zr[c] - mean) * rstd : zr[c];
            const float u = norm ? xh * gamma[c] + beta[c] : xh;
            float du = dyr[c];
            if (drop_p > 0.f) du *= drop_scale(sd, (unsigned long long)row * C + c, drop_p, inv_keep);
            if (act == 1 && !(u > 0.f)) du = 0.f;
            return du;
        };
        float s1 = 0.f, s2 = 0.f;
        if (norm) {
            for (int c = lane; c < C; c += 32) {
                float xh;
                const float du = du_at(c, xh);
                atomicAdd(&sacc[c], du * xh);
                atomicAdd(&sacc[C + c], du);
                const float dxh = du * gamma[c];
                s1 += dxh; s2 += dxh * xh;
            }
            s1 = warp_sum(s1) * invC; s2 = warp_sum(s2) * invC;
        }
        for (int c = lane; c < C; c += 32) {
            float xh;
            const float du = du_at(c, xh);
            const float d = norm ? rstd * (du * gamma[c] - s1 - xh * s2) : du;
            if (dz_hi) st_split1(dz_hi, dz_lo, row * ldp + c, d);
            else dz[row * lddz + c] = d;
            atomicAdd(&sacc[2 * C + c], d);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (norm) { atomicAdd(dgamma + i, sacc[i]); atomicAdd(dbeta + i, sacc[C + i]); }
        if (dbias) atomicAdd(dbias + i, sacc[2 * C + i]);
    }
}

// backward of the highway tail: dz [rows][2C] and the residual-path gradient dxres = do*(1-g)
__global__ void hc_post_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ z, long long ldz,
                                   const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
                                   const float* __restrict__ g1, const float* __restrict__ b1,
                                   const float* __restrict__ g2, const float* __restrict__ b2,
                                   float* __restrict__ dz, long long lddz, float* __restrict__ dxres, long long lddx,
                                   float* __restrict__ dg1, float* __restrict__ db1, float* __restrict__ dg2,
                                   float* __restrict__ db2, float* __restrict__ dbias,
                                   int rows, int C, int norm, float drop_p, unsigned long long seed, const long long* step,
                                   LccGate lcc) {
    pdl_grid_sync();
    extern __shared__ float sacc[];            // [6][C]: dg1, db1, dg2, db2, dbias(H1), dbias(H2)
    for (int i = threadIdx.x; i < 6 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float invC = 1.f / (float)C;
    const unsigned long long sd = eff_seed(seed, step);
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const float* z1 = z + row * ldz;
        const float* z2 = z1 + C;
        const float* xr = x + row * ldx;
        const float* dyr = dy + row * lddy;
        float* dz1 = dz + row * lddz;
        float* dz2 = dz1 + C;
        float m1 = 0.f, r1 = 1.f, m2 = 0.f, r2 = 1.f;
        if (norm) { m1 = stats[row * 4]; r1 = stats[row * 4 + 1]; m2 = stats[row * 4 + 2]; r2 = stats[row * 4 + 3]; }
        float a1 = 0.f, a2 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float xh1 = norm ? (z1[c] - m1) * r1 : z1[c];
            const float xh2 = norm ? (z2[c] - m2) * r2 : z2[c];
            const float u1 = norm ? xh1 * g1[c] + b1[c] : xh1;
            const float h = norm ? xh2 * g2[c] + b2[c] : xh2;
            const float g = sigmoidf_(u1);
            float d_o = dyr[c];
            if (drop_p > 0.f) d_o *= drop_scale(sd, (unsigned long long)row * C + c, drop_p, inv_keep);
            dxres[row * lddx + c] = d_o * (1.f - g);
            float gl = 1.f;
            if (lcc.table) {           // out = g * (gl * h) + (1 - g) * x;  d out / d table = d_o * g * h * gl (1 - gl), summed over time
                gl = lcc_gate(lcc, row, c, C);
                lcc.scratch[row * C + c] = d_o * g * h * gl * (1.f - gl);
            }
            const float du1 = d_o * (gl * h - xr[c]) * g * (1.f - g);
            const float du2 = d_o * g * gl;
            if (norm) {
                atomicAdd(&sacc[c], du1 * xh1); atomicAdd(&sacc[C + c], du1);
                atomicAdd(&sacc[2 * C + c], du2 * xh2); atomicAdd(&sacc[3 * C + c], du2);
                const float e1 = du1 * g1[c], e2 = du2 * g2[c];
                a1 += e1; a2 += e1 * xh1; c1 += e2; c2 += e2 * xh2;
                dz1[c] = e1; dz2[c] = e2;
            } else {
                dz1[c] = du1; dz2[c] = du2;
                atomicAdd(&sacc[4 * C + c], du1); atomicAdd(&sacc[5 * C + c], du2);
            }
        }
        if (norm) {
            a1 = warp_sum(a1) * invC; a2 = warp_sum(a2) * invC; c1 = warp_sum(c1) * invC; c2 = warp_sum(c2) * invC;
            for (int c = lane; c < C; c += 32) {
                const float xh1 = (z1[c] - m1) * r1, xh2 = (z2[c] - m2) * r2;
                const float d1 = r1 * (dz1[c] - a1 - xh1 * a2);
                const float d2 = r2 * (dz2[c] - c1 - xh2 * c2);
                dz1[c] = d1; dz2[c] = d2;
                atomicAdd(&sacc[4 * C + c], d1); atomicAdd(&sacc[5 * C + c], d2);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (norm) {
            atomicAdd(dg1 + i, sacc[i]); atomicAdd(db1 + i, sacc[C + i]);
            atomicAdd(dg2 + i, sacc[2 * C + i]); atomicAdd(db2 + i, sacc[3 * C + i]);
        }
        if (dbias) { atomicAdd(dbias + i, sacc[4 * C + i]); atomicAdd(dbias + C + i, sacc[5 * C + i]); }
    }
}

// ------------------------------------------------------------------------------------------------ attention
__device__ __forceinline__ float guide_w(int n, int t, float inv_maxN, float inv_maxT, float inv_2g2) {
    const float d = (float)t * inv_maxT - (float)n * inv_maxN;           // utils.py:155-161
    return 1.f - __expf(-d * d * inv_2g2);
}
// Per-utterance guides from the batch (hp.attention_guide_dir, architectures.py:57-58): gw[b][n][t] inside the
// batch-padded [Ng, Tg] block, `pad` outside it (1.0 for the guided loss, 0.0 for the MSE variant; :263, :275).
struct GuideTensor {
    const float* w; long long item_stride; long long ld; int Ng, Tg; float pad; int mse;
    // gradient of the CDP / Ain / Aout terms (architectures.py:283-321):  dA[b][t][n] += col_g[b][n] + (col_h[b][n] + c_aout) * log A + c_aout
    const float* col_g; const float* col_h; float c_aout;
};
__device__ __forceinline__ float guide_t(const GuideTensor& G, int b, int n, int t) {
    return (n < G.Ng && t < G.Tg) ? __ldg(G.w + (long long)b * G.item_stride + (long long)n * G.ld + t) : G.pad;
}

// in: S[b][t][0..N) scaled scores.  out: probabilities in place, optional transposed alignments [B][N][T],
// argmax (first maximum), guided-attention partial sum  sum_{n<maxN,t<maxT} A*W  (architectures.py:258-270)
__global__ void softmax_fwd_kernel(float* __restrict__ S, long long ldS, int B, int T, int N,
                                   const int* __restrict__ prev_max, int win,
                                   float* __restrict__ align_t, int* __restrict__ argmax_out,
                                   double* __restrict__ att_acc, int maxN, int maxT, float g, GuideTensor G,
                                   unsigned short* __restrict__ p_hi, unsigned short* __restrict__ p_lo, long long ldp) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float inv_maxN = 1.f / (float)maxN, inv_maxT = 1.f / (float)maxT, inv_2g2 = 1.f / (2.f * g * g);
    float att_part = 0.f;
    const long long rows = (long long)B * T;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const int b = (int)(row / T), t = (int)(row - (long long)b * T);
        float* sr = S + row * ldS;
        int lo = 0, hi = N;
        if (prev_max) {
            if (win > 0) { lo = prev_max[b]; hi = lo + win; }     // allowed keys: lo <= n < hi (networks.py:304-306)
            else hi = prev_max[b];                                // win <= 0: per-item key count, n < hi (networks.py:308-309)
        }
        float mx = -INFINITY; int arg = 0x7fffffff;
        for (int n = lane; n < N; n += 32) {
            float v = sr[n];
            if (n < lo || n >= hi) v = ATT_MASK_VALUE;
            if (v > mx) { mx = v; arg = n; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float sum = 0.f;
        for (int n = lane; n < N; n += 32) {
            float v = sr[n];
            if (n < lo || n >= hi) v = ATT_MASK_VALUE;
            const float e = __expf(v - mx);
            sr[n] = e; sum += e;
        }
        const float inv = 1.f / warp_sum(sum);
        for (int n = lane; n < N; n += 32) {
            const float a = sr[n] * inv;
            sr[n] = a;
            if (p_hi) st_split1(p_hi, p_lo, row * ldp + n, a);      // operand planes of the A.V / A^T.dR products
            if (align_t) align_t[((long long)b * N + n) * T + t] = a;
            if (att_acc && n < maxN && t < maxT) {
                if (!G.w) att_part += a * guide_w(n, t, inv_maxN, inv_maxT, inv_2g2);
                else {
                    const float w = guide_t(G, b, n, t);
                    att_part += G.mse ? (a - w) * (a - w) : a * w;       // sum (A - W)^2  |  sum |A * W|  (A, W >= 0)
                }
            }
        }
        if (argmax_out && lane == 0) argmax_out[row] = arg;
    }
    if (att_acc) {
        __shared__ float red[32];
        att_part = warp_sum(att_part);
        if (lane == 0) red[threadIdx.x >> 5] = att_part;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < wpb ? red[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(att_acc, (double)v);
        }
    }
}

// The guided-attention sum of softmax_fwd_kernel for alignments that were not computed here (FixedAttention,
// networks.py:327-358: the externally supplied duration matrix is reported through the same loss term):
// att_acc += sum_{n < maxN, t < maxT} A*W  (or (A - W)^2 for the MSE variant)
__global__ void guide_sum_kernel(const float* __restrict__ A, long long ldA, int B, int T, int N, double* __restrict__ att_acc,
                                 int maxN, int maxT, float g, GuideTensor G) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float inv_maxN = 1.f / (float)maxN, inv_maxT = 1.f / (float)maxT, inv_2g2 = 1.f / (2.f * g * g);
    float part = 0.f;
    const long long rows = (long long)B * T;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const int b = (int)(row / T), t = (int)(row - (long long)b * T);
        if (t >= maxT) continue;
        for (int n = lane; n < N && n < maxN; n += 32) {
            const float a = A[row * ldA + n];
            if (!G.w) part += a * guide_w(n, t, inv_maxN, inv_maxT, inv_2g2);
            else {
                const float w = guide_t(G, b, n, t);
                part += G.mse ? (a - w) * (a - w) : a * w;
            }
        }
    }
    __shared__ float red[32];
    part = warp_sum(part);
    if (lane == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < wpb ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(att_acc, (double)v);
    }
}

// dA (in) -> dS (in place):  dS = A * (dA' - sum_n A*dA'),  dA' = dA + att_coef * W[n][t]   (MSE variant: + att_coef * 2 (A - W))
__global__ void softmax_bwd_kernel(const float* __restrict__ A, long long ldA, float* __restrict__ dA, long long lddA,
                                   int B, int T, int N, float att_coef, int maxN, int maxT, float g, GuideTensor G,
                                   unsigned short* __restrict__ ds_hi, unsigned short* __restrict__ ds_lo, long long ldp) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float inv_maxN = 1.f / (float)maxN, inv_maxT = 1.f / (float)maxT, inv_2g2 = 1.f / (2.f * g * g);
    const long long rows = (long long)B * T;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const int b = (int)(row / T), t = (int)(row - (long long)b * T);
        const float* ar = A + row * ldA;
        float* dr = dA + row * lddA;
        float dot = 0.f;
        for (int n = lane; n < N; n += 32) {
            float d = dr[n];
            if (att_coef != 0.f && n < maxN && t < maxT) {
                if (!G.w) d += att_coef * guide_w(n, t, inv_maxN, inv_maxT, inv_2g2);
                else {
                    const float w = guide_t(G, b, n, t);
                    d += att_coef * (G.mse ? 2.f * (ar[n] - w) : w);
                }
            }
            if (G.col_g) {
                const float a = ar[n];
                if (a > 0.f) d += G.col_g[(long long)b * N + n] + (G.col_h[(long long)b * N + n] + G.c_aout) * __logf(a) + G.c_aout;
                else d += G.col_g[(long long)b * N + n];
            }
            dr[n] = d;
            dot += ar[n] * d;
        }
        dot = warp_sum(dot);
        for (int n = lane; n < N; n += 32) {
            const float ds = ar[n] * (dr[n] - dot);
            dr[n] = ds;
            if (ds_hi) st_split1(ds_hi, ds_lo, row * ldp + n, ds);
        }
    }
}

// "Confidence through attention" terms (architectures.py:283-321) over the alignments of the batch, A[b][t][n]:
//   s[b][n] = sum_t A,  e[b][n] = sum_t A log A  (0 log 0 = 0)
//   CDP  = sum_{b,n} log(1 + (1 - s)^2) / (B N)
//   Ain  = - sum_{b,n} sum_t P log P / (B N log T),  P = A / s  ->  sum_t P log P = e / s - log s
//   Aout = - sum_{b,t} sum_n A log A / (B T log N)   (rows of A already sum to one)  = - sum_{b,n} e / (B T log N)
// One block per (b, 32 keys): lanes along n (coalesced), 8 warps stride over t.  Adds the three sums to acc3 and stores
// the per-key factors of the gradient, d/dA[b][t][n] = col_g + col_h log A (+ the Aout part, which needs no statistics):
//   col_g = c_cdp * (-2 (1 - s) / (1 + (1 - s)^2)) - c_ain * e / s^2,   col_h = c_ain / s      (zero where s == 0)
__global__ void att_extra_kernel(const float* __restrict__ A, long long ldA, int B, int T, int N, float c_cdp, float c_ain,
                                 float* __restrict__ col_g, float* __restrict__ col_h, double* __restrict__ acc3) {
    pdl_grid_sync();
    __shared__ float ss[8][32], se[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, n = blockIdx.x * 32 + lane;
    float s = 0.f, e = 0.f;
    if (n < N) {
        const float* ap = A + (long long)b * T * ldA + n;
        for (int t = warp; t < T; t += 8) {
            const float a = ap[(long long)t * ldA];
            s += a;
            if (a > 0.f) e += a * __logf(a);
        }
    }
    ss[warp][lane] = s; se[warp][lane] = e;
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < 8; ++w) { s += ss[w][lane]; e += se[w][lane]; }
        float cdp = 0.f, ain = 0.f, aout = 0.f;
        if (n < N) {
            const float r = 1.f - s;
            cdp = __logf(1.f + r * r);
            float g = c_cdp * (-2.f * r / (1.f + r * r)), h = 0.f;
            if (s > 0.f) {
                ain = e / s - __logf(s);
                g -= c_ain * e / (s * s);
                h = c_ain / s;
            }
            aout = e;
            col_g[(long long)b * N + n] = g;
            col_h[(long long)b * N + n] = h;
        }
        cdp = warp_sum(cdp); ain = warp_sum(ain); aout = warp_sum(aout);
        if (lane == 0) { atomicAdd(acc3, (double)cdp); atomicAdd(acc3 + 1, (double)ain); atomicAdd(acc3 + 2, (double)aout); }
    }
}

// loss_components[5..7] = CDP, Ain, Aout; the weighted terms join the total only under the legacy lw_* pattern
// (architectures.py:333-349: the loss_weights dict branch does not add them)
__global__ void att_extra_finalize_kernel(const double* __restrict__ acc3, float* __restrict__ comps, double n_keys,
                                          double n_frames, double logT, double logN, float w_cdp, float w_ain, float w_aout,
                                          int add_to_total) {
    pdl_grid_sync();
    const double cdp = acc3[0] / n_keys, ain = -acc3[1] / n_keys / logT, aout = -acc3[2] / n_frames / logN;
    comps[5] = (float)cdp; comps[6] = (float)ain; comps[7] = (float)aout;
    if (add_to_total) {
        double tot = comps[0];
        if (w_cdp != 0.f) tot += w_cdp * cdp;
        if (w_ain != 0.f) tot += w_ain * ain;
        if (w_aout != 0.f) tot += w_aout * aout;
        comps[0] = (float)tot;
    }
}

// fp32 rows -> split-bf16 planes (hi = bf16(x), lo = bf16(x - hi)): operands that reach a GEMM from outside the
// row-wise kernels (fed K / V at synthesis, the decoder-input gradient) get their copy-engine format here
__global__ void split_planes_kernel(const float* __restrict__ x, long long ldx, long long rows, int C,
                                    unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, long long ldp) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const float* xr = x + row * ldx;
        if (!(C & 3) && !(ldx & 3) && !(ldp & 3)) {
            for (int c = lane * 4; c < C; c += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
                uint2 hh, ll;
                split4(v, hh, ll);
                *reinterpret_cast<uint2*>(hi + row * ldp + c) = hh;
                *reinterpret_cast<uint2*>(lo + row * ldp + c) = ll;
            }
        } else {
            for (int c = lane; c < C; c += 32) st_split1(hi, lo, row * ldp + c, xr[c]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ embedding
// modules.py:38-42: row 0 of the table reads as zeros and receives no gradient.
__global__ void embed_fwd_kernel(const int* __restrict__ ids, const float* __restrict__ table, float* __restrict__ out,
                                 long long ldo, int rows, int E) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const int id = ids[row];
        for (int c = lane; c < E; c += 32) out[row * ldo + c] = id == 0 ? 0.f : table[(long long)id * E + c];
    }
}
__global__ void embed_bwd_kernel(const int* __restrict__ ids, const float* __restrict__ dout, long long ldo,
                                 float* __restrict__ dtable, int rows, int E) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const int id = ids[row];
        if (id == 0) continue;
        for (int c = lane; c < E; c += 32) atomicAdd(dtable + (long long)id * E + c, dout[row * ldo + c]);
    }
}

// ------------------------------------------------------------------------------------------------ losses
// architectures.py:147-170 / :245-255.  acc[0..2] += sum|Y-t|, sum BCE(logits,t), sum (Y-t)^2.
// dlogits = (w_l1*sign(Y-t)*Y' + w_bd*(Y-t) + w_l2*2(Y-t)*Y') / n   with Y' = Y(1-Y) (1 if not squashed)
__global__ void recon_loss_kernel(const float* __restrict__ logits, long long ldl, const float* __restrict__ target,
                                  long long ldt, float* __restrict__ dlogits, long long ldd,
                                  long long rows, int C, int squash, float w_l1, float w_bd, float w_l2,
                                  double* __restrict__ acc) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float inv_n = 1.f / ((float)rows * (float)C);
    float s1 = 0.f, sb = 0.f, s2 = 0.f;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        for (int c = lane; c < C; c += 32) {
            const float x = logits[row * ldl + c], t = target[row * ldt + c];
            const float y = squash ? sigmoidf_(x) : x;
            const float d = y - t;
            s1 += fabsf(d); s2 += d * d;
            if (squash) sb += fmaxf(x, 0.f) - x * t + log1pf(__expf(-fabsf(x)));
            if (dlogits) {
                const float yp = squash ? y * (1.f - y) : 1.f;
                const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                float gsum = w_l1 * sg * yp + w_l2 * 2.f * d * yp;
                if (squash) gsum += w_bd * d;
                dlogits[row * ldd + c] = gsum * inv_n;
            }
        }
    }
    __shared__ float red[3][32];
    s1 = warp_sum(s1); sb = warp_sum(sb); s2 = warp_sum(s2);
    if (lane == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = sb; red[2][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 96) {
        const int k = threadIdx.x >> 5;
        float v = lane < wpb ? red[k][lane] : 0.f;
        v = warp_sum(v);
        if (lane == 0) atomicAdd(acc + k, (double)v);
    }
}

// loss_components (architectures.py:173, :352-355): out = [loss, L1, BD, (att,) L2]
__global__ void loss_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out, double n_recon, double n_att,
                                     float w_l1, float w_bd, float w_att, float w_l2, int has_att, int squash) {
    pdl_grid_sync();
    const double l1 = acc[0] / n_recon, bd = squash ? acc[1] / n_recon : 0.0, l2 = acc[2] / n_recon;
    const double att = has_att ? acc[3] / n_att : 0.0;
    const double loss = w_l1 * l1 + w_bd * bd + w_att * att + w_l2 * l2;
    out[0] = (float)loss; out[1] = (float)l1; out[2] = (float)bd;
    if (has_att) { out[3] = (float)att; out[4] = (float)l2; } else { out[3] = (float)l2; }
}

// ------------------------------------------------------------------------------------------------ optimiser
// architectures.py:101-128 + utils.py:167-170: Noam lr on step+1, TF Adam bias correction folded into lr_t.
__global__ void adam_prepare_kernel(const long long* __restrict__ global_step, float* __restrict__ lr_t, float lr0,
                                    float beta1, float beta2, int decay_lr, float warmup) {
    pdl_grid_sync();
    const double t = (double)(*global_step + 1);
    double lr = lr0;
    if (decay_lr) lr = (double)lr0 * sqrt((double)warmup) * fmin(t * pow((double)warmup, -1.5), 1.0 / sqrt(t));
    lr_t[0] = (float)(lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
    lr_t[1] = (float)lr;
}
__global__ void step_inc_kernel(long long* global_step) {
    pdl_grid_sync(); *global_step += 1; }

// clip_by_value(g*grad_scale, +-clip) -> m,v update -> theta -= lr_t * m / (sqrt(v) + eps)   (TF epsilon placement)
__global__ void adam_clip_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                 const float* __restrict__ g, long long n, const float* __restrict__ lr_t,
                                 float beta1, float beta2, float eps, float clip, float grad_scale) {
    pdl_grid_sync();
    const float lr = lr_t[0];
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pg = reinterpret_cast<const float4*>(g)[i];
        float4 pp = reinterpret_cast<float4*>(p)[i], pm = reinterpret_cast<float4*>(m)[i], pv = reinterpret_cast<float4*>(v)[i];
        float* gg = &pg.x; float* ppp = &pp.x; float* mm = &pm.x; float* vv = &pv.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gr = fminf(fmaxf(gg[e] * grad_scale, -clip), clip);
            mm[e] = beta1 * mm[e] + (1.f - beta1) * gr;
            vv[e] = beta2 * vv[e] + (1.f - beta2) * gr * gr;
            ppp[e] -= lr * mm[e] / (sqrtf(vv[e]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = pm; reinterpret_cast<float4*>(v)[i] = pv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gr = fminf(fmaxf(g[i] * grad_scale, -clip), clip);
        const float mi = beta1 * m[i] + (1.f - beta1) * gr;
        const float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
        m[i] = mi; v[i] = vi;
        p[i] -= lr * mi / (sqrtf(vi) + eps);
    }
}

}  // namespace oph

// =================================================================================================
// Vectorised fast paths for C = 128*VEC (256, 512, 1024): each lane owns 4*VEC channels as float4s, the row lives
// in registers (one HBM read per element), per-channel sums of the backward pass accumulate in registers across
// the rows a warp owns and are flushed once per warp.
namespace oph {

template <int VEC>
__device__ __forceinline__ void ld_row(const float* __restrict__ p, int lane, float4 (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(p + i * 128 + lane * 4));
}
template <int VEC>
__device__ __forceinline__ void st_row(float* __restrict__ p, int lane, const float4 (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) *reinterpret_cast<float4*>(p + i * 128 + lane * 4) = v[i];
}
// same row as split-bf16 planes (hi = bf16(x), lo = bf16(x - hi)): the operand format the GEMM producers copy verbatim
template <int VEC>
__device__ __forceinline__ void st_row_planes(unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int lane,
                                              const float4 (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i].x, v[i].y), h1 = __floats2bfloat162_rn(v[i].z, v[i].w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v[i].x - f0.x, v[i].y - f0.y), l1 = __floats2bfloat162_rn(v[i].z - f1.x, v[i].w - f1.y);
        uint2 hh, ll;
        hh.x = *reinterpret_cast<const uint32_t*>(&h0); hh.y = *reinterpret_cast<const uint32_t*>(&h1);
        ll.x = *reinterpret_cast<const uint32_t*>(&l0); ll.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(hi + i * 128 + lane * 4) = hh;
        *reinterpret_cast<uint2*>(lo + i * 128 + lane * 4) = ll;
    }
}
template <int VEC>
__device__ __forceinline__ RowStats reg_stats(const float4 (&v)[VEC]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.f / (128.f * VEC));
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float var = warp_sum(q) * (1.f / (128.f * VEC));
    return {mean, rsqrtf(var + LN_EPS)};
}
#define OPH_F4(v, e) (reinterpret_cast<float*>(&(v))[e])

template <int VEC>
__global__ void __launch_bounds__(256) hc_post_fwd_vec_kernel(
        const float* __restrict__ z, long long ldz, const float* __restrict__ x, long long ldx,
        const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
        const float* __restrict__ b2, float* __restrict__ y, long long ldy, unsigned short* __restrict__ y_hi,
        unsigned short* __restrict__ y_lo, long long ldp, float* __restrict__ stats,
        int rows, int norm, float drop_p, unsigned long long seed, const long long* step) {
    pdl_grid_sync();
    constexpr int C = 128 * VEC;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned long long sd = eff_seed(seed, step);
    float4 G1[VEC], B1[VEC], G2[VEC], B2[VEC];
    if (norm) { ld_row<VEC>(g1, lane, G1); ld_row<VEC>(b1, lane, B1); ld_row<VEC>(g2, lane, G2); ld_row<VEC>(b2, lane, B2); }
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        float4 z1[VEC], z2[VEC], xv[VEC], o[VEC];
        ld_row<VEC>(z + row * ldz, lane, z1);
        ld_row<VEC>(z + row * ldz + C, lane, z2);
        ld_row<VEC>(x + row * ldx, lane, xv);
        RowStats s1 = {0.f, 1.f}, s2 = {0.f, 1.f};
        if (norm) { s1 = reg_stats<VEC>(z1); s2 = reg_stats<VEC>(z2); }
        if (stats && lane == 0) *reinterpret_cast<float4*>(stats + row * 4) = make_float4(s1.mean, s1.rstd, s2.mean, s2.rstd);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float u1 = OPH_F4(z1[i], e), u2 = OPH_F4(z2[i], e);
                if (norm) {
                    u1 = (u1 - s1.mean) * s1.rstd * OPH_F4(G1[i], e) + OPH_F4(B1[i], e);
                    u2 = (u2 - s2.mean) * s2.rstd * OPH_F4(G2[i], e) + OPH_F4(B2[i], e);
                }
                const float g = sigmoidf_(u1);
                float r = g * u2 + (1.f - g) * OPH_F4(xv[i], e);
                if (drop_p > 0.f) r *= drop_scale(sd, (unsigned long long)row * C + i * 128 + lane * 4 + e, drop_p, inv_keep);
                OPH_F4(o[i], e) = r;
            }
        st_row<VEC>(y + row * ldy, lane, o);
        if (y_hi) st_row_planes<VEC>(y_hi + row * ldp, y_lo + row * ldp, lane, o);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) ln_act_fwd_vec_kernel(
        const float* __restrict__ z, long long ldz, const float* __restrict__ gamma, const float* __restrict__ beta,
        float* __restrict__ y, long long ldy, float* __restrict__ y_sig, long long ldys,
        unsigned short* __restrict__ y_hi, unsigned short* __restrict__ y_lo, long long ldp, float* __restrict__ stats,
        int rows, int act, int norm, float drop_p, unsigned long long seed, const long long* step) {
    pdl_grid_sync();
    constexpr int C = 128 * VEC;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned long long sd = eff_seed(seed, step);
    float4 G[VEC], Bt[VEC];
    if (norm) { ld_row<VEC>(gamma, lane, G); ld_row<VEC>(beta, lane, Bt); }
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        float4 zv[VEC], o[VEC], sg[VEC];
        ld_row<VEC>(z + row * ldz, lane, zv);
        RowStats st = {0.f, 1.f};
        if (norm) st = reg_stats<VEC>(zv);
        if (stats && lane == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(st.mean, st.rstd);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float u = OPH_F4(zv[i], e);
                if (norm) u = (u - st.mean) * st.rstd * OPH_F4(G[i], e) + OPH_F4(Bt[i], e);
                if (y_sig) OPH_F4(sg[i], e) = sigmoidf_(u);
                float a = act == 1 ? fmaxf(u, 0.f) : u;
                if (drop_p > 0.f) a *= drop_scale(sd, (unsigned long long)row * C + i * 128 + lane * 4 + e, drop_p, inv_keep);
                OPH_F4(o[i], e) = a;
            }
        st_row<VEC>(y + row * ldy, lane, o);
        if (y_hi) st_row_planes<VEC>(y_hi + row * ldp, y_lo + row * ldp, lane, o);
        if (y_sig) st_row<VEC>(y_sig + row * ldys, lane, sg);
    }
}

// flush per-lane register accumulators: shared atomics (one per lane and channel per warp), then global atomics
template <int VEC, int NACC>
__device__ __forceinline__ void flush_acc(float (&acc)[NACC][VEC * 4], float* sacc, int lane, float* const (&dst)[NACC]) {
    constexpr int C = 128 * VEC;
#pragma unroll
    for (int k = 0; k < NACC; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(&sacc[k * C + i * 128 + lane * 4 + e], acc[k][i * 4 + e]);
    __syncthreads();
    for (int idx = threadIdx.x; idx < NACC * C; idx += blockDim.x) {
        float* d = dst[idx / C];
        if (d) atomicAdd(d + (idx % C), sacc[idx]);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) ln_act_bwd_vec_kernel(
        const float* __restrict__ dy, long long lddy, const float* __restrict__ z, long long ldz,
        const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
        float* __restrict__ dz, long long lddz, unsigned short* __restrict__ dz_hi, unsigned short* __restrict__ dz_lo,
        long long ldp, float* __restrict__ dgamma, float* __restrict__ dbeta,
        float* __restrict__ dbias, int rows, int act, int norm, float drop_p, unsigned long long seed,
        const long long* step) {
    pdl_grid_sync();
    constexpr int C = 128 * VEC;
    extern __shared__ float sacc[];            // [3][C]
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float invC = 1.f / (float)C;
    const unsigned long long sd = eff_seed(seed, step);
    float4 G[VEC], Bt[VEC];
    if (norm) { ld_row<VEC>(gamma, lane, G); ld_row<VEC>(beta, lane, Bt); }
    float acc[3][VEC * 4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < VEC * 4; ++i) acc[k][i] = 0.f;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        float4 zv[VEC], dv[VEC], ev[VEC];
        ld_row<VEC>(z + row * ldz, lane, zv);
        ld_row<VEC>(dy + row * lddy, lane, dv);
        float mean = 0.f, rstd = 1.f;
        if (norm) { const float2 s = __ldg(reinterpret_cast<const float2*>(stats + row * 2)); mean = s.x; rstd = s.y; }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xh = norm ? (OPH_F4(zv[i], e) - mean) * rstd : OPH_F4(zv[i], e);
                const float u = norm ? xh * OPH_F4(G[i], e) + OPH_F4(Bt[i], e) : xh;
                float du = OPH_F4(dv[i], e);
                if (drop_p > 0.f) du *= drop_scale(sd, (unsigned long long)row * C + i * 128 + lane * 4 + e, drop_p, inv_keep);
                if (act == 1 && !(u > 0.f)) du = 0.f;
                if (norm) {
                    acc[0][i * 4 + e] += du * xh; acc[1][i * 4 + e] += du;
                    const float dxh = du * OPH_F4(G[i], e);
                    s1 += dxh; s2 += dxh * xh;
                    OPH_F4(ev[i], e) = dxh; OPH_F4(zv[i], e) = xh;
                } else {
                    OPH_F4(ev[i], e) = du; acc[2][i * 4 + e] += du;
                }
            }
        if (norm) {
            s1 = warp_sum(s1) * invC; s2 = warp_sum(s2) * invC;
#pragma unroll
            for (int i = 0; i < VEC; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = rstd * (OPH_F4(ev[i], e) - s1 - OPH_F4(zv[i], e) * s2);
                    OPH_F4(ev[i], e) = d; acc[2][i * 4 + e] += d;
                }
        }
        if (dz_hi) st_row_planes<VEC>(dz_hi + row * ldp, dz_lo + row * ldp, lane, ev);
        else st_row<VEC>(dz + row * lddz, lane, ev);
    }
    float* const dst[3] = {norm ? dgamma : nullptr, norm ? dbeta : nullptr, dbias};
    flush_acc<VEC, 3>(acc, sacc, lane, dst);
}

}  // namespace oph

// =================================================================================================
// Backward of the highway tail, bandwidth-oriented layout: WPR warps share one row (256 channels per warp, two
// float4 per lane), so the per-lane register accumulators of the six per-channel sums stay at 48 registers for every
// width (C = 256 * WPR).  Each warp streams its rows through a private ring of `depth` shared-memory slots filled
// with 16-byte async copies (LDGSTS), several rows ahead of the one being processed: one 8-warp block per SM keeps
// >100 KB of loads in flight.  gamma / beta vectors live in shared memory; the four per-row LN-backward sums are
// combined across the warps of a row through shared memory and a named barrier.
namespace oph {

constexpr int HCB_SLOT = 8 * 512 + 32;                  // 8 float4 per lane + the row's (mean, rstd) x 2

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int WPR>
__global__ void __launch_bounds__(256, 2) hc_post_bwd_wide_kernel(
        const float* __restrict__ dy, long long lddy, const float* __restrict__ z, long long ldz,
        const float* __restrict__ x, long long ldx, const float* __restrict__ stats,
        const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
        const float* __restrict__ b2, unsigned short* __restrict__ dz_hi, unsigned short* __restrict__ dz_lo,
        long long ldp, float* __restrict__ dxres, long long lddx,
        float* __restrict__ dg1, float* __restrict__ db1, float* __restrict__ dg2, float* __restrict__ db2,
        float* __restrict__ dbias, int rows, float drop_p, unsigned long long seed, const long long* step, int depth) {
    pdl_grid_sync();
    constexpr int C = 256 * WPR;
    constexpr int GROUPS = 8 / WPR;                      // rows in flight per block
    extern __shared__ __align__(16) float smem_f[];
    float* sacc = smem_f;                               // [6][C]
    float* spar = smem_f + 6 * C;                       // [4][C]: g1, b1, g2, b2
    float* sx = spar + 4 * C;                           // [2 parities][8 warps][4] partial row sums
    uint8_t* ring = reinterpret_cast<uint8_t*>(sx + 64);   // [8 warps][depth][HCB_SLOT]
    for (int i = threadIdx.x; i < 6 * C; i += 256) sacc[i] = 0.f;
    for (int i = threadIdx.x; i < C; i += 256) { spar[i] = g1[i]; spar[C + i] = b1[i]; spar[2 * C + i] = g2[i]; spar[3 * C + i] = b2[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / WPR, part = warp % WPR;
    const int cbase = part * 256 + lane * 4;            // first channel of this lane's float4 i is cbase + 128 * i
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float invC = 1.f / (float)C;
    const unsigned long long sd = eff_seed(seed, step);
    float acc[6][8];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;

    const long long stride = (long long)gridDim.x * GROUPS;
    long long row = (long long)blockIdx.x * GROUPS + grp;
    uint8_t* myring = ring + (size_t)warp * depth * HCB_SLOT;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(myring) + lane * 16;
    auto issue = [&](long long r, int slot) {           // this lane's 9 x 16 bytes of row r -> ring slot
        if (r < rows) {
            const uint32_t d = ring_u32 + slot * HCB_SLOT;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                cp_async_16(d + (0 + i) * 512, z + r * ldz + cbase + 128 * i);
                cp_async_16(d + (2 + i) * 512, z + r * ldz + C + cbase + 128 * i);
                cp_async_16(d + (4 + i) * 512, x + r * ldx + cbase + 128 * i);
                cp_async_16(d + (6 + i) * 512, dy + r * lddy + cbase + 128 * i);
            }
            if (lane == 0) cp_async_16(d + 8 * 512, stats + r * 4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int d = 0; d < depth - 1; ++d) issue(row + d * stride, d);
    int it = 0, slot = 0;
    for (; row < rows; row += stride, ++it) {
        {
            int ns = slot + depth - 1; if (ns >= depth) ns -= depth;
            issue(row + (long long)(depth - 1) * stride, ns);
        }
        // all but the newest depth-1 groups are complete: the slot of this row has landed
        if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (depth == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
        else if (depth == 5) asm volatile("cp.async.wait_group 4;" ::: "memory");
        else asm volatile("cp.async.wait_group 5;" ::: "memory");
        __syncwarp();                                    // lane 0's copy of the row statistics is read by every lane
        const uint8_t* sl = myring + slot * HCB_SLOT + lane * 16;
        float4 z1[2], z2[2], xv[2], dv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            z1[i] = *reinterpret_cast<const float4*>(sl + (0 + i) * 512);
            z2[i] = *reinterpret_cast<const float4*>(sl + (2 + i) * 512);
            xv[i] = *reinterpret_cast<const float4*>(sl + (4 + i) * 512);
            dv[i] = *reinterpret_cast<const float4*>(sl + (6 + i) * 512);
        }
        const float4 st = *reinterpret_cast<const float4*>(myring + slot * HCB_SLOT + 8 * 512);
        if (++slot == depth) slot = 0;
        const float m1 = st.x, r1 = st.y, m2 = st.z, r2 = st.w;
        float a1 = 0.f, a2 = 0.f, c1 = 0.f, c2 = 0.f;
        float4 e1v[2], e2v[2], xr[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 G1 = *reinterpret_cast<const float4*>(spar + cbase + 128 * i);
            float4 B1 = *reinterpret_cast<const float4*>(spar + C + cbase + 128 * i);
            float4 G2 = *reinterpret_cast<const float4*>(spar + 2 * C + cbase + 128 * i);
            float4 B2 = *reinterpret_cast<const float4*>(spar + 3 * C + cbase + 128 * i);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xh1 = (OPH_F4(z1[i], e) - m1) * r1;
                const float xh2 = (OPH_F4(z2[i], e) - m2) * r2;
                const float u1 = xh1 * OPH_F4(G1, e) + OPH_F4(B1, e);
                const float h = xh2 * OPH_F4(G2, e) + OPH_F4(B2, e);
                const float g = sigmoidf_(u1);
                float d_o = OPH_F4(dv[i], e);
                if (drop_p > 0.f) d_o *= drop_scale(sd, (unsigned long long)row * C + cbase + 128 * i + e, drop_p, inv_keep);
                OPH_F4(xr[i], e) = d_o * (1.f - g);
                const float du1 = d_o * (h - OPH_F4(xv[i], e)) * g * (1.f - g);
                const float du2 = d_o * g;
                acc[0][i * 4 + e] += du1 * xh1; acc[1][i * 4 + e] += du1;
                acc[2][i * 4 + e] += du2 * xh2; acc[3][i * 4 + e] += du2;
                const float e1 = du1 * OPH_F4(G1, e), e2 = du2 * OPH_F4(G2, e);
                a1 += e1; a2 += e1 * xh1; c1 += e2; c2 += e2 * xh2;
                OPH_F4(e1v[i], e) = e1; OPH_F4(e2v[i], e) = e2;
                OPH_F4(z1[i], e) = xh1; OPH_F4(z2[i], e) = xh2;
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(dxres + row * lddx + cbase + 128 * i) = xr[i];
        a1 = warp_sum(a1); a2 = warp_sum(a2); c1 = warp_sum(c1); c2 = warp_sum(c2);
        if (WPR > 1) {                                   // combine the partial sums of the warps sharing this row
            float* my = sx + ((it & 1) * 8 + warp) * 4;
            if (lane == 0) *reinterpret_cast<float4*>(my) = make_float4(a1, a2, c1, c2);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
            a1 = a2 = c1 = c2 = 0.f;
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                const float4 o = *reinterpret_cast<const float4*>(sx + ((it & 1) * 8 + grp * WPR + w) * 4);
                a1 += o.x; a2 += o.y; c1 += o.z; c2 += o.w;
            }
        }
        a1 *= invC; a2 *= invC; c1 *= invC; c2 *= invC;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d1 = r1 * (OPH_F4(e1v[i], e) - a1 - OPH_F4(z1[i], e) * a2);
                const float d2 = r2 * (OPH_F4(e2v[i], e) - c1 - OPH_F4(z2[i], e) * c2);
                OPH_F4(e1v[i], e) = d1; OPH_F4(e2v[i], e) = d2;
                acc[4][i * 4 + e] += d1; acc[5][i * 4 + e] += d2;
            }
        {   // dz in the GEMMs' operand format: split-bf16 planes [rows][2C]
            unsigned short* ph = dz_hi + row * ldp + cbase;
            unsigned short* pl = dz_lo + row * ldp + cbase;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                uint2 hh, ll;
                split4(e1v[i], hh, ll);
                *reinterpret_cast<uint2*>(ph + 128 * i) = hh; *reinterpret_cast<uint2*>(pl + 128 * i) = ll;
                split4(e2v[i], hh, ll);
                *reinterpret_cast<uint2*>(ph + C + 128 * i) = hh; *reinterpret_cast<uint2*>(pl + C + 128 * i) = ll;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // flush: per-lane sums -> shared (one atomic per warp and channel) -> global (one atomic per block and channel)
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(&sacc[k * C + cbase + 128 * i + e], acc[k][i * 4 + e]);
    __syncthreads();
    float* const dst[6] = {dg1, db1, dg2, db2, dbias, dbias ? dbias + C : nullptr};
    for (int idx = threadIdx.x; idx < 6 * C; idx += 256) {
        float* d = dst[idx / C];
        if (d) atomicAdd(d + (idx % C), sacc[idx]);
    }
}

}  // namespace oph

// =================================================================================================
// Forward highway tail in the same streaming layout as hc_post_bwd_wide_kernel: WPR warps per row, gamma / beta in
// shared memory (loaded once per block, not once per warp), rows prefetched through per-warp cp.async rings.  Row
// moments: each warp takes the two-pass mean / M2 of its 256 channels, the warps of a row merge them exactly
// (equal-sized groups: M2 = sum M2_w + 256 * sum (mean_w - mean)^2).
namespace oph {

constexpr int HCF_SLOT = 6 * 512;                       // z1, z2, x: two float4 per lane each

template <int WPR>
__global__ void __launch_bounds__(256) hc_post_fwd_wide_kernel(
        const float* __restrict__ z, long long ldz, const float* __restrict__ x, long long ldx,
        const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
        const float* __restrict__ b2, float* __restrict__ y, long long ldy, unsigned short* __restrict__ y_hi,
        unsigned short* __restrict__ y_lo, long long ldp, float* __restrict__ stats,
        int rows, float drop_p, unsigned long long seed, const long long* step, int depth, long long row_base) {
    // row_base: index of row 0 within the layer's activation (the dropout mask is a function of the absolute element index)
    pdl_grid_sync();
    constexpr int C = 256 * WPR;
    constexpr int GROUPS = 8 / WPR;
    extern __shared__ __align__(16) float smem_f[];
    float* spar = smem_f;                               // [4][C]: g1, b1, g2, b2
    float* sx = spar + 4 * C;                           // [2 parities][8 warps][4] partial moments
    uint8_t* ring = reinterpret_cast<uint8_t*>(sx + 64);
    for (int i = threadIdx.x; i < C; i += 256) { spar[i] = g1[i]; spar[C + i] = b1[i]; spar[2 * C + i] = g2[i]; spar[3 * C + i] = b2[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / WPR, part = warp % WPR;
    const int cbase = part * 256 + lane * 4;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned long long sd = eff_seed(seed, step);
    const long long stride = (long long)gridDim.x * GROUPS;
    long long row = (long long)blockIdx.x * GROUPS + grp;
    uint8_t* myring = ring + (size_t)warp * depth * HCF_SLOT;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(myring) + lane * 16;
    auto issue = [&](long long r, int slot) {
        if (r < rows) {
            const uint32_t d = ring_u32 + slot * HCF_SLOT;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                cp_async_16(d + (0 + i) * 512, z + r * ldz + cbase + 128 * i);
                cp_async_16(d + (2 + i) * 512, z + r * ldz + C + cbase + 128 * i);
                cp_async_16(d + (4 + i) * 512, x + r * ldx + cbase + 128 * i);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int d = 0; d < depth - 1; ++d) issue(row + d * stride, d);
    int it = 0, slot = 0;
    for (; row < rows; row += stride, ++it) {
        {
            int ns = slot + depth - 1; if (ns >= depth) ns -= depth;
            issue(row + (long long)(depth - 1) * stride, ns);
        }
        if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else asm volatile("cp.async.wait_group 3;" ::: "memory");
        const uint8_t* sl = myring + slot * HCF_SLOT + lane * 16;
        float4 z1[2], z2[2], xv[2], o[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            z1[i] = *reinterpret_cast<const float4*>(sl + (0 + i) * 512);
            z2[i] = *reinterpret_cast<const float4*>(sl + (2 + i) * 512);
            xv[i] = *reinterpret_cast<const float4*>(sl + (4 + i) * 512);
        }
        if (++slot == depth) slot = 0;
        // two-pass moments of this warp's 256 channels
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) { s1 += (z1[i].x + z1[i].y) + (z1[i].z + z1[i].w); s2 += (z2[i].x + z2[i].y) + (z2[i].z + z2[i].w); }
        float m1 = warp_sum(s1) * (1.f / 256.f), m2 = warp_sum(s2) * (1.f / 256.f);
        float q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = OPH_F4(z1[i], e) - m1, b = OPH_F4(z2[i], e) - m2;
                q1 += a * a; q2 += b * b;
            }
        q1 = warp_sum(q1); q2 = warp_sum(q2);
        if (WPR > 1) {                                   // merge the equal-sized groups of the warps sharing this row
            float* my = sx + ((it & 1) * 8 + warp) * 4;
            if (lane == 0) *reinterpret_cast<float4*>(my) = make_float4(m1, q1, m2, q2);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
            float mm1 = 0.f, mm2 = 0.f;
            float4 ow[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) { ow[w] = *reinterpret_cast<const float4*>(sx + ((it & 1) * 8 + grp * WPR + w) * 4); mm1 += ow[w].x; mm2 += ow[w].z; }
            mm1 *= (1.f / WPR); mm2 *= (1.f / WPR);
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                t1 += ow[w].y + 256.f * (ow[w].x - mm1) * (ow[w].x - mm1);
                t2 += ow[w].w + 256.f * (ow[w].z - mm2) * (ow[w].z - mm2);
            }
            m1 = mm1; m2 = mm2; q1 = t1; q2 = t2;
        }
        const float r1 = rsqrtf(q1 * (1.f / (float)C) + LN_EPS), r2 = rsqrtf(q2 * (1.f / (float)C) + LN_EPS);
        if (stats && lane == 0 && part == 0) *reinterpret_cast<float4*>(stats + row * 4) = make_float4(m1, r1, m2, r2);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 G1 = *reinterpret_cast<const float4*>(spar + cbase + 128 * i);
            float4 B1 = *reinterpret_cast<const float4*>(spar + C + cbase + 128 * i);
            float4 G2 = *reinterpret_cast<const float4*>(spar + 2 * C + cbase + 128 * i);
            float4 B2 = *reinterpret_cast<const float4*>(spar + 3 * C + cbase + 128 * i);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float u1 = (OPH_F4(z1[i], e) - m1) * r1 * OPH_F4(G1, e) + OPH_F4(B1, e);
                const float u2 = (OPH_F4(z2[i], e) - m2) * r2 * OPH_F4(G2, e) + OPH_F4(B2, e);
                const float g = sigmoidf_(u1);
                float r = g * u2 + (1.f - g) * OPH_F4(xv[i], e);
                if (drop_p > 0.f) r *= drop_scale(sd, (unsigned long long)(row_base + row) * C + cbase + 128 * i + e, drop_p, inv_keep);
                OPH_F4(o[i], e) = r;
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            *reinterpret_cast<float4*>(y + row * ldy + cbase + 128 * i) = o[i];
            if (y_hi) {
                uint2 hh, ll;
                split4(o[i], hh, ll);
                *reinterpret_cast<uint2*>(y_hi + row * ldp + cbase + 128 * i) = hh;
                *reinterpret_cast<uint2*>(y_lo + row * ldp + cbase + 128 * i) = ll;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace oph

// =================================================================================================
// Backward of the conv tail (LayerNorm + ReLU + dropout) in the streaming layout of hc_post_bwd_wide_kernel: WPR warps
// per row, gamma / beta in shared memory, rows prefetched through per-warp cp.async rings, three per-channel sums
// (dgamma, dbeta, dbias) in 24 registers per lane for every width C = 256 * WPR.
namespace oph {

constexpr int LNB_SLOT = 4 * 512 + 32;                  // z, dy: two float4 per lane each + (mean, rstd)

template <int WPR>
__global__ void __launch_bounds__(256) ln_act_bwd_wide_kernel(
        const float* __restrict__ dy, long long lddy, const float* __restrict__ z, long long ldz,
        const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
        unsigned short* __restrict__ dz_hi, unsigned short* __restrict__ dz_lo, long long ldp,
        float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
        int rows, int act, float drop_p, unsigned long long seed, const long long* step, int depth) {
    pdl_grid_sync();
    constexpr int C = 256 * WPR;
    constexpr int GROUPS = 8 / WPR;
    extern __shared__ __align__(16) float smem_f[];
    float* sacc = smem_f;                               // [3][C]
    float* spar = smem_f + 3 * C;                       // [2][C]: gamma, beta
    float* sx = spar + 2 * C;                           // [2 parities][8 warps][2] partial row sums
    uint8_t* ring = reinterpret_cast<uint8_t*>(sx + 32);
    for (int i = threadIdx.x; i < 3 * C; i += 256) sacc[i] = 0.f;
    for (int i = threadIdx.x; i < C; i += 256) { spar[i] = gamma[i]; spar[C + i] = beta[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / WPR, part = warp % WPR;
    const int cbase = part * 256 + lane * 4;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float invC = 1.f / (float)C;
    const unsigned long long sd = eff_seed(seed, step);
    float acc[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    const long long stride = (long long)gridDim.x * GROUPS;
    long long row = (long long)blockIdx.x * GROUPS + grp;
    uint8_t* myring = ring + (size_t)warp * depth * LNB_SLOT;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(myring) + lane * 16;
    auto issue = [&](long long r, int slot) {
        if (r < rows) {
            const uint32_t d = ring_u32 + slot * LNB_SLOT;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                cp_async_16(d + (0 + i) * 512, z + r * ldz + cbase + 128 * i);
                cp_async_16(d + (2 + i) * 512, dy + r * lddy + cbase + 128 * i);
            }
            if (lane == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d - lane * 16 + 4 * 512), "l"(stats + r * 2) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int d = 0; d < depth - 1; ++d) issue(row + d * stride, d);
    int it = 0, slot = 0;
    for (; row < rows; row += stride, ++it) {
        {
            int ns = slot + depth - 1; if (ns >= depth) ns -= depth;
            issue(row + (long long)(depth - 1) * stride, ns);
        }
        if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else asm volatile("cp.async.wait_group 3;" ::: "memory");
        __syncwarp();
        const uint8_t* sl = myring + slot * LNB_SLOT + lane * 16;
        float4 zv[2], dv[2], ev[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            zv[i] = *reinterpret_cast<const float4*>(sl + (0 + i) * 512);
            dv[i] = *reinterpret_cast<const float4*>(sl + (2 + i) * 512);
        }
        const float2 st = *reinterpret_cast<const float2*>(myring + slot * LNB_SLOT + 4 * 512);
        if (++slot == depth) slot = 0;
        const float mean = st.x, rstd = st.y;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 G = *reinterpret_cast<const float4*>(spar + cbase + 128 * i);
            float4 Bt = *reinterpret_cast<const float4*>(spar + C + cbase + 128 * i);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xh = (OPH_F4(zv[i], e) - mean) * rstd;
                const float u = xh * OPH_F4(G, e) + OPH_F4(Bt, e);
                float du = OPH_F4(dv[i], e);
                if (drop_p > 0.f) du *= drop_scale(sd, (unsigned long long)row * C + cbase + 128 * i + e, drop_p, inv_keep);
                if (act == 1 && !(u > 0.f)) du = 0.f;
                acc[0][i * 4 + e] += du * xh; acc[1][i * 4 + e] += du;
                const float dxh = du * OPH_F4(G, e);
                s1 += dxh; s2 += dxh * xh;
                OPH_F4(ev[i], e) = dxh; OPH_F4(zv[i], e) = xh;
            }
        }
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (WPR > 1) {
            float* my = sx + ((it & 1) * 8 + warp) * 2;
            if (lane == 0) *reinterpret_cast<float2*>(my) = make_float2(s1, s2);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
            s1 = s2 = 0.f;
#pragma unroll
            for (int w = 0; w < WPR; ++w) {
                const float2 o = *reinterpret_cast<const float2*>(sx + ((it & 1) * 8 + grp * WPR + w) * 2);
                s1 += o.x; s2 += o.y;
            }
        }
        s1 *= invC; s2 *= invC;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = rstd * (OPH_F4(ev[i], e) - s1 - OPH_F4(zv[i], e) * s2);
                OPH_F4(ev[i], e) = d; acc[2][i * 4 + e] += d;
            }
            uint2 hh, ll;
            split4(ev[i], hh, ll);
            *reinterpret_cast<uint2*>(dz_hi + row * ldp + cbase + 128 * i) = hh;
            *reinterpret_cast<uint2*>(dz_lo + row * ldp + cbase + 128 * i) = ll;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(&sacc[k * C + cbase + 128 * i + e], acc[k][i * 4 + e]);
    __syncthreads();
    float* const dst[3] = {dgamma, dbeta, dbias};
    for (int idx = threadIdx.x; idx < 3 * C; idx += 256) {
        float* d = dst[idx / C];
        if (d) atomicAdd(d + (idx % C), sacc[idx]);
    }
}

}  // namespace oph

// =================================================================================================
// Conv tails for channel counts that are not 256 / 512 / 1024 (80 mel bins, 513 / 1025 magnitude bins: the output layers
// of AudioDec and SSRN, which run on the largest activations of the model).  Same arithmetic as ln_act_fwd / bwd_kernel,
// but 16-byte accesses: rows are padded to a multiple of 4 floats, a group of WPR warps owns a row, lane l of part p holds
// the float4s (i * WPR + p) * 32 + l of it in registers (i < MAXV), elements >= C are masked out of every sum.  Each lane
// always sees the same channels, so the per-channel sums of the backward pass stay in registers until the end.
namespace oph {

template <int MAXV, int WPR>
__global__ void __launch_bounds__(256) ln_act_fwd_any_kernel(
        const float* __restrict__ z, long long ldz, const float* __restrict__ gamma, const float* __restrict__ beta,
        float* __restrict__ y, long long ldy, float* __restrict__ y_sig, long long ldys,
        unsigned short* __restrict__ y_hi, unsigned short* __restrict__ y_lo, long long ldp, float* __restrict__ stats,
        int rows, int C, int act, int norm, float drop_p, unsigned long long seed, const long long* step) {
    pdl_grid_sync();
    constexpr int GROUPS = 8 / WPR;
    extern __shared__ __align__(16) float smem_f[];
    const int Cp = (C + 3) & ~3;
    float* spar = smem_f;                               // [2][Cp]: gamma, beta (zero beyond C)
    float* sx = smem_f + 2 * Cp;                        // [2 parities][8 warps][2]
    for (int i = threadIdx.x; i < Cp; i += 256) { spar[i] = (norm && i < C) ? gamma[i] : 0.f; spar[Cp + i] = (norm && i < C) ? beta[i] : 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / WPR, part = warp % WPR;
    const int nv = Cp >> 2;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float invC = 1.f / (float)C;
    const unsigned long long sd = eff_seed(seed, step);
    int it = 0;
    for (long long row = (long long)blockIdx.x * GROUPS + grp; row < rows; row += (long long)gridDim.x * GROUPS, ++it) {
        float4 v[MAXV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = (i * WPR + part) * 32 + lane;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < nv) {
                v[i] = __ldg(reinterpret_cast<const float4*>(z + row * ldz) + j);
#pragma unroll
                for (int e = 0; e < 4; ++e) if (j * 4 + e >= C) OPH_F4(v[i], e) = 0.f;
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        float mean = 0.f, rstd = 1.f;
        if (norm) {
            s = warp_sum(s);
            if (WPR > 1) {
                float* my = sx + ((it & 1) * 8 + warp) * 2;
                if (lane == 0) my[0] = s;
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
                s = 0.f;
#pragma unroll
                for (int w = 0; w < WPR; ++w) s += sx[((it & 1) * 8 + grp * WPR + w) * 2];
            }
            mean = s * invC;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                const int j = (i * WPR + part) * 32 + lane;
                if (j < nv) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (j * 4 + e < C) { const float d = OPH_F4(v[i], e) - mean; q += d * d; }
                }
            }
            q = warp_sum(q);
            if (WPR > 1) {
                float* my = sx + ((it & 1) * 8 + warp) * 2;
                if (lane == 0) my[1] = q;
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
                q = 0.f;
#pragma unroll
                for (int w = 0; w < WPR; ++w) q += sx[((it & 1) * 8 + grp * WPR + w) * 2 + 1];
            }
            rstd = rsqrtf(q * invC + LN_EPS);
        }
        if (stats && lane == 0 && part == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = (i * WPR + part) * 32 + lane;
            if (j >= nv) continue;
            const int c0 = j * 4;
            const float4 G = *reinterpret_cast<const float4*>(spar + c0), Bt = *reinterpret_cast<const float4*>(spar + Cp + c0);
            float4 o, sg;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float u = OPH_F4(v[i], e);
                if (norm) u = (u - mean) * rstd * reinterpret_cast<const float*>(&G)[e] + reinterpret_cast<const float*>(&Bt)[e];
                OPH_F4(sg, e) = sigmoidf_(u);
                float a = act == 1 ? fmaxf(u, 0.f) : u;
                if (drop_p > 0.f) a *= drop_scale(sd, (unsigned long long)row * C + c0 + e, drop_p, inv_keep);
                OPH_F4(o, e) = a;
            }
            if (c0 + 4 <= C) {
                *reinterpret_cast<float4*>(y + row * ldy + c0) = o;
                if (y_sig) *reinterpret_cast<float4*>(y_sig + row * ldys + c0) = sg;
                if (y_hi) {
                    uint2 hh, ll;
                    split4(o, hh, ll);
                    *reinterpret_cast<uint2*>(y_hi + row * ldp + c0) = hh;
                    *reinterpret_cast<uint2*>(y_lo + row * ldp + c0) = ll;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (c0 + e >= C) break;
                    y[row * ldy + c0 + e] = OPH_F4(o, e);
                    if (y_sig) y_sig[row * ldys + c0 + e] = OPH_F4(sg, e);
                    if (y_hi) st_split1(y_hi, y_lo, row * ldp + c0 + e, OPH_F4(o, e));
                }
            }
        }
    }
}

// backward: dz as split-bf16 planes [rows][ldp] (or fp32 when dz_hi == NULL), dgamma / dbeta / dbias accumulated into
template <int MAXV, int WPR>
__global__ void __launch_bounds__(256, 3) ln_act_bwd_any_kernel(
        const float* __restrict__ dy, long long lddy, const float* __restrict__ z, long long ldz,
        const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
        float* __restrict__ dz, long long lddz, unsigned short* __restrict__ dz_hi, unsigned short* __restrict__ dz_lo,
        long long ldp, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
        int rows, int C, int act, int norm, float drop_p, unsigned long long seed, const long long* step) {
    pdl_grid_sync();
    constexpr int GROUPS = 8 / WPR;
    extern __shared__ __align__(16) float smem_f[];
    const int Cp = (C + 3) & ~3;
    float* spar = smem_f;                               // [2][Cp]
    float* sx = smem_f + 2 * Cp;                        // [2 parities][8 warps][2]
    for (int i = threadIdx.x; i < Cp; i += 256) { spar[i] = (norm && i < C) ? gamma[i] : 0.f; spar[Cp + i] = (norm && i < C) ? beta[i] : 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / WPR, part = warp % WPR;
    const int nv = Cp >> 2;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float invC = 1.f / (float)C;
    const unsigned long long sd = eff_seed(seed, step);
    float acc[3][MAXV * 4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < MAXV * 4; ++i) acc[k][i] = 0.f;
    int it = 0;
    for (long long row = (long long)blockIdx.x * GROUPS + grp; row < rows; row += (long long)gridDim.x * GROUPS, ++it) {
        float mean = 0.f, rstd = 1.f;
        if (norm) { const float2 st = __ldg(reinterpret_cast<const float2*>(stats + row * 2)); mean = st.x; rstd = st.y; }
        float4 xh[MAXV], ev[MAXV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = (i * WPR + part) * 32 + lane;
            xh[i] = ev[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j >= nv) continue;
            const int c0 = j * 4;
            const float4 zv = __ldg(reinterpret_cast<const float4*>(z + row * ldz) + j);
            const float4 dv = __ldg(reinterpret_cast<const float4*>(dy + row * lddy) + j);
            const float4 G = *reinterpret_cast<const float4*>(spar + c0), Bt = *reinterpret_cast<const float4*>(spar + Cp + c0);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (c0 + e >= C) continue;
                const float zz = reinterpret_cast<const float*>(&zv)[e];
                const float x_ = norm ? (zz - mean) * rstd : zz;
                const float u = norm ? x_ * reinterpret_cast<const float*>(&G)[e] + reinterpret_cast<const float*>(&Bt)[e] : x_;
                float du = reinterpret_cast<const float*>(&dv)[e];
                if (drop_p > 0.f) du *= drop_scale(sd, (unsigned long long)row * C + c0 + e, drop_p, inv_keep);
                if (act == 1 && !(u > 0.f)) du = 0.f;
                if (norm) {
                    acc[0][i * 4 + e] += du * x_; acc[1][i * 4 + e] += du;
                    const float dxh = du * reinterpret_cast<const float*>(&G)[e];
                    s1 += dxh; s2 += dxh * x_;
                    OPH_F4(ev[i], e) = dxh; OPH_F4(xh[i], e) = x_;
                } else {
                    OPH_F4(ev[i], e) = du; acc[2][i * 4 + e] += du;
                }
            }
        }
        if (norm) {
            s1 = warp_sum(s1); s2 = warp_sum(s2);
            if (WPR > 1) {
                float* my = sx + ((it & 1) * 8 + warp) * 2;
                if (lane == 0) *reinterpret_cast<float2*>(my) = make_float2(s1, s2);
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPR * 32) : "memory");
                s1 = s2 = 0.f;
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    const float2 o = *reinterpret_cast<const float2*>(sx + ((it & 1) * 8 + grp * WPR + w) * 2);
                    s1 += o.x; s2 += o.y;
                }
            }
            s1 *= invC; s2 *= invC;
        }
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = (i * WPR + part) * 32 + lane;
            if (j >= nv) continue;
            const int c0 = j * 4;
            if (norm) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (c0 + e >= C) continue;
                    const float d = rstd * (OPH_F4(ev[i], e) - s1 - OPH_F4(xh[i], e) * s2);
                    OPH_F4(ev[i], e) = d; acc[2][i * 4 + e] += d;
                }
            }
            if (dz_hi) {                                 // plane rows are padded to 8 elements: whole float4s fit
                uint2 hh, ll;
                split4(ev[i], hh, ll);
                *reinterpret_cast<uint2*>(dz_hi + row * ldp + c0) = hh;
                *reinterpret_cast<uint2*>(dz_lo + row * ldp + c0) = ll;
            } else if (c0 + 4 <= C) {
                *reinterpret_cast<float4*>(dz + row * lddz + c0) = ev[i];
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (c0 + e < C) dz[row * lddz + c0 + e] = OPH_F4(ev[i], e);
            }
        }
    }
    // flush: per-block sums in shared memory first, then one global atomic per channel and block (a global atomic per
    // lane made the 80-channel output layer's backward 87 us long: 835 k atomics on 240 addresses)
    float* sacc = smem_f + 2 * Cp + 32;                 // [3][Cp]
    for (int i = threadIdx.x; i < 3 * Cp; i += 256) sacc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < 2 && !norm) continue;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = (i * WPR + part) * 32 + lane;
            if (j >= nv) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(&sacc[k * Cp + j * 4 + e], acc[k][i * 4 + e]);
        }
    }
    __syncthreads();
    float* const dst[3] = {norm ? dgamma : nullptr, norm ? dbeta : nullptr, dbias};
    for (int idx = threadIdx.x; idx < 3 * Cp; idx += 256) {
        const int k = idx / Cp, c = idx - k * Cp;
        if (dst[k] && c < C) atomicAdd(dst[k] + c, sacc[idx]);
    }
}

}  // namespace oph


// =================================================================================================
// Per-speaker channel gates behind a conv1d layer (modules.py:141-144): out = gate[b][c] * y0, gate = sigmoid(embed(code_b)).
namespace oph {

__global__ void lcc_fwd_kernel(const float* __restrict__ y0, long long ld0, LccGate lcc, float* __restrict__ out, long long ldo,
                               float* __restrict__ out_sig, long long lds, unsigned short* __restrict__ o_hi,
                               unsigned short* __restrict__ o_lo, long long ldp, long long rows, int C) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        for (int c = lane; c < C; c += 32) {
            const float v = y0[row * ld0 + c] * lcc_gate(lcc, row, c, C);
            out[row * ldo + c] = v;
            if (out_sig) out_sig[row * lds + c] = sigmoidf_(v);
            if (o_hi) st_split1(o_hi, o_lo, row * ldp + c, v);
        }
    }
}
// dy0 = gate * dy;  scratch = dy * y0 * gate (1 - gate)  (summed over time per item by lcc_reduce_kernel)
__global__ void lcc_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ y0, long long ld0,
                               LccGate lcc, float* __restrict__ dy0, long long ldd0, long long rows, int C) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        for (int c = lane; c < C; c += 32) {
            const float gl = lcc_gate(lcc, row, c, C), d = dy[row * lddy + c];
            dy0[row * ldd0 + c] = d * gl;
            lcc.scratch[row * C + c] = d * y0[row * ld0 + c] * gl * (1.f - gl);
        }
    }
}
// dtable[code_b][c] += sum_t scratch[b][t][c]   (block = (32 channels, item b); the zero-pad row 0 gets no gradient)
__global__ void lcc_reduce_kernel(const float* __restrict__ scratch, const int* __restrict__ codes, float* __restrict__ dtable,
                                  int L, int C) {
    pdl_grid_sync();
    __shared__ float ss[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, c = blockIdx.x * 32 + lane;
    const int code = codes[b];
    float s = 0.f;
    if (c < C && code != 0)
        for (int t = warp; t < L; t += 8) s += scratch[((long long)b * L + t) * C + c];
    ss[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < C && code != 0) {
        for (int w = 1; w < 8; ++w) s += ss[w][lane];
        atomicAdd(dtable + (long long)code * C + c, s);
    }
}

}  // namespace oph
