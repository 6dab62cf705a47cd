// networks.Attention forward (networks.py:286-325) as ONE kernel: for a 128-row query tile of one utterance
//     S = Q K^T / sqrt(d)  ->  window mask  ->  softmax over the keys  ->  argmax, guided-attention loss partial,
//     alignments  ->  R = P V
// with both contractions on tcgen05 (3-term split-bf16, fp32 accumulation in tensor memory) and nothing of S or P going
// through global memory between them: S lives in tensor-memory columns [0, 256), the probabilities are split into bf16
// (hi, lo) planes straight into shared memory in the K-major SWIZZLE_128B layout the second product consumes, R
// accumulates in tensor-memory columns [256, 512).  A (fp32 and / or planes) and the transposed alignments are written
// only when the caller asks for them (training saves A for the backward pass; synthesis fetches the alignments).
//
// One CTA (cta_group::1) per work unit (utterance b, 128 query rows), persistent over the units.  Warp roles: 16 softmax /
// epilogue warps (lane = query row of TMEM lane quadrant warp % 4; warp / 4 selects 64 of the up to 256 key columns), one
// copy-engine thread, one MMA-issuing thread.  Shared memory is one 224 KiB arena used twice per unit:
//   phase 1: two stages of {Q k-block (hi | lo, 32 KiB), K k-block (hi | lo, up to 64 KiB)} over d in blocks of 64
//   phase 2-3: P (hi | lo per 64-key block, 32 KiB each) followed by one or two 64 KiB slots for V blocks (64 keys x d).
// Needs d == 256 and N <= 256 (the dc_tts shapes: d = 256, max_N = 150 / 180).
#pragma once
#include <cuda.h>
#include "oph_ptx.cuh"

namespace oph {

struct AttnArgs {
    int B, T, N, d;
    int Npad;                  // N rounded up to 16: the N extent of the first product / K extent (in 16s) of the second
    int nblk;                  // 64-key blocks of the second product
    int vslots;                // 1 or 2 V slots in the arena
    float scale;               // 1 / sqrt(d)
    const int* prev_max; int win;
    float* A; long long ldA;                        // [B][T][ldA] probabilities (nullable)
    unsigned short* Ahi; unsigned short* Alo; long long ldAp;   // their planes (nullable)
    float* align_t;            // [B][N][T] (nullable)
    int* argmax;               // [B][T] (nullable)
    double* att_acc; int maxN, maxT; float g; GuideTensor G;
    float* R; long long ldr;                        // [B][T][ldr]
    unsigned short* Rhi; unsigned short* Rlo; long long ldrp;   // planes of R (nullable)
    alignas(64) CUtensorMap tmQ_hi;
    alignas(64) CUtensorMap tmQ_lo;
    alignas(64) CUtensorMap tmK_hi;
    alignas(64) CUtensorMap tmK_lo;
    alignas(64) CUtensorMap tmV_hi;
    alignas(64) CUtensorMap tmV_lo;
};

constexpr int ATT_BM = 128;
constexpr int ATT_STAGE = 96 * 1024;            // Q k-block 32 KiB + K k-block up to 64 KiB
constexpr int ATT_ARENA = 224 * 1024;
constexpr int ATT_PBLK = 32 * 1024;             // P block: hi 16 KiB | lo 16 KiB
constexpr int ATT_VSLOT = 64 * 1024;            // V block: hi 32 KiB | lo 32 KiB (64 keys x 256 channels)
constexpr int ATT_SW = 16;                       // softmax / epilogue warps
constexpr int ATT_THREADS = (ATT_SW + 2) * 32;
constexpr int ATT_SMEM = ATT_ARENA + 128 + 2048;  // arena (1024-byte aligned base), barriers, row exchange
// barriers
constexpr int AB_FULL = 0;      // [2] phase-1 stage loaded
constexpr int AB_EMPTY = 2;     // [2] phase-1 stage consumed
constexpr int AB_SFULL = 4;     // S complete in tensor memory
constexpr int AB_PREADY = 5;    // P planes in shared memory (one arrival per softmax warp)
constexpr int AB_VFULL = 6;     // [2]
constexpr int AB_VEMPTY = 8;    // [2]
constexpr int AB_RFULL = 10;    // R complete
constexpr int AB_REMPTY = 11;   // R drained (one arrival per softmax warp)
constexpr int AB_NUM = 12;

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_fused_kernel(const __grid_constant__ AttnArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* arena = smem_raw;
    if (smem_u32(smem_raw) & 1023u) __trap();                   // SWIZZLE_128B tiles need the 1024-byte aligned base
    uint64_t* bars = reinterpret_cast<uint64_t*>(arena + ATT_ARENA);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AB_NUM);
    float* xch = reinterpret_cast<float*>(arena + ATT_ARENA + 128);     // [4 column quarters][128 rows]: max, argmax, sum in turn
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(AB_FULL + i), 1); mbar_init(BAR(AB_EMPTY + i), 1); mbar_init(BAR(AB_VFULL + i), 1); mbar_init(BAR(AB_VEMPTY + i), 1); }
        mbar_init(BAR(AB_SFULL), 1); mbar_init(BAR(AB_PREADY), ATT_SW); mbar_init(BAR(AB_RFULL), 1); mbar_init(BAR(AB_REMPTY), ATT_SW);
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == ATT_SW + 1) tmem_alloc1<512>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_grid_sync();

    const int tiles = (p.T + ATT_BM - 1) / ATT_BM;
    const int units = p.B * tiles;
    const int KB1 = p.d / 64;                                   // k-blocks of the first product (4)
    const uint32_t arena_u32 = smem_u32(arena);
    const uint32_t v_base = arena_u32 + p.nblk * ATT_PBLK;      // V slots follow the P blocks
    const uint32_t k_bytes = (uint32_t)p.Npad * 128u;           // one plane of a K k-block

    if (warp == ATT_SW) {
        // ================================================================ copy engine (one thread)
        if (lane == 0) {
            tma_prefetch_desc(&p.tmQ_hi); tma_prefetch_desc(&p.tmQ_lo); tma_prefetch_desc(&p.tmK_hi);
            tma_prefetch_desc(&p.tmK_lo); tma_prefetch_desc(&p.tmV_hi); tma_prefetch_desc(&p.tmV_lo);
            int st = 0, st_par = 1, vs = 0, vs_par = 1, rf_par = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int b = u / tiles, t0 = (u - b * tiles) * ATT_BM;
                for (int kb = 0; kb < KB1; ++kb) {
                    mbar_wait(BAR(AB_EMPTY + st), st_par);
                    const uint32_t dst = arena_u32 + st * ATT_STAGE;
                    mbar_arrive_expect_tx(BAR(AB_FULL + st), 2 * 16384 + 2 * k_bytes);
                    tma_load_3d(dst, &p.tmQ_hi, kb * 64, t0, b, BAR(AB_FULL + st));
                    tma_load_3d(dst + 16384, &p.tmQ_lo, kb * 64, t0, b, BAR(AB_FULL + st));
                    tma_load_3d(dst + 32768, &p.tmK_hi, kb * 64, 0, b, BAR(AB_FULL + st));
                    tma_load_3d(dst + 65536, &p.tmK_lo, kb * 64, 0, b, BAR(AB_FULL + st));
                    if (++st == 2) { st = 0; st_par ^= 1; }
                }
                // the V slots overlap the phase-1 stages: wait until the first product has consumed them
                mbar_wait(BAR(AB_SFULL), rf_par);
                for (int blk = 0; blk < p.nblk; ++blk) {
                    mbar_wait(BAR(AB_VEMPTY + vs), vs_par);
                    const uint32_t dst = v_base + vs * ATT_VSLOT;
                    mbar_arrive_expect_tx(BAR(AB_VFULL + vs), ATT_VSLOT);
                    for (int c = 0; c < 4; ++c) {               // 64-channel boxes of 64 keys each
                        tma_load_3d(dst + c * 8192, &p.tmV_hi, c * 64, blk * 64, b, BAR(AB_VFULL + vs));
                        tma_load_3d(dst + 32768 + c * 8192, &p.tmV_lo, c * 64, blk * 64, b, BAR(AB_VFULL + vs));
                    }
                    if (++vs == p.vslots) { vs = 0; vs_par ^= 1; }
                }
                // the next unit's phase-1 stages overlap P and V: wait for the second product
                mbar_wait(BAR(AB_RFULL), rf_par);
                rf_par ^= 1;
            }
        }
        __syncwarp();
    } else if (warp == ATT_SW + 1) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc1 = make_idesc_bf16(ATT_BM, p.Npad, 0, 0);
            const uint32_t idesc2 = make_idesc_bf16(ATT_BM, 256, 0, 1);
            int st = 0, st_par = 0, vs = 0, vs_par = 0, u_par = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                // ---- S = Q K^T (the softmax warps of the previous unit are done with S: they arrived on R_EMPTY after it)
                for (int kb = 0; kb < KB1; ++kb) {
                    mbar_wait(BAR(AB_FULL + st), st_par);
                    tc_fence_after();
                    const uint32_t q_hi = arena_u32 + st * ATT_STAGE, q_lo = q_hi + 16384, k_hi = q_hi + 32768, k_lo = q_hi + 65536;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t dqh = make_sdesc(q_hi + ks * 32, 16, 1024), dql = make_sdesc(q_lo + ks * 32, 16, 1024);
                        const uint64_t dkh = make_sdesc(k_hi + ks * 32, 16, 1024), dkl = make_sdesc(k_lo + ks * 32, 16, 1024);
                        umma1_bf16(tmem_base, dqh, dkh, idesc1, (kb | ks) != 0);
                        umma1_bf16(tmem_base, dqh, dkl, idesc1, 1);
                        umma1_bf16(tmem_base, dql, dkh, idesc1, 1);
                    }
                    umma1_commit(BAR(AB_EMPTY + st));
                    if (++st == 2) { st = 0; st_par ^= 1; }
                }
                umma1_commit(BAR(AB_SFULL));
                // ---- R = P V
                mbar_wait(BAR(AB_PREADY), u_par);
                mbar_wait(BAR(AB_REMPTY), u_par ^ 1);            // R columns drained by the previous unit's epilogue
                tc_fence_after();
                const int ksteps = p.Npad / 16;
                for (int blk = 0; blk < p.nblk; ++blk) {
                    mbar_wait(BAR(AB_VFULL + vs), vs_par);
                    tc_fence_after();
                    const uint32_t p_hi = arena_u32 + blk * ATT_PBLK, p_lo = p_hi + 16384;
                    const uint32_t v_hi = v_base + vs * ATT_VSLOT, v_lo = v_hi + 32768;
                    const int nks = min(4, ksteps - blk * 4);
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t dph = make_sdesc(p_hi + ks * 32, 16, 1024), dpl = make_sdesc(p_lo + ks * 32, 16, 1024);
                        const uint64_t dvh = make_sdesc(v_hi + ks * 2048, 8192, 1024), dvl = make_sdesc(v_lo + ks * 2048, 8192, 1024);
                        umma1_bf16(tmem_base + 256, dph, dvh, idesc2, (blk | ks) != 0);
                        umma1_bf16(tmem_base + 256, dph, dvl, idesc2, 1);
                        umma1_bf16(tmem_base + 256, dpl, dvh, idesc2, 1);
                    }
                    umma1_commit(BAR(AB_VEMPTY + vs));
                    if (++vs == p.vslots) { vs = 0; vs_par ^= 1; }
                }
                umma1_commit(BAR(AB_RFULL));
                u_par ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ================================================================ softmax + epilogue warps
        const int q = warp & 3, half = warp >> 2;                // TMEM lane quadrant, key-column quarter (0..3)
        const int r = q * 32 + lane;                             // query row within the tile = TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float inv_maxN = 1.f / (float)p.maxN, inv_maxT = 1.f / (float)p.maxT, inv_2g2 = 1.f / (2.f * p.g * p.g);
        const int cbeg = half * 64, cend = min(p.Npad, cbeg + 64);       // this warp's S columns
        float att_part = 0.f;
        int u_par = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int b = u / tiles, t0 = (u - b * tiles) * ATT_BM;
            const int t = t0 + r;
            const bool row_ok = t < p.T;
            int lo = 0, hi = p.N;
            if (p.prev_max) {
                if (p.win > 0) { lo = p.prev_max[b]; hi = lo + p.win; }
                else hi = p.prev_max[b];
            }
            mbar_wait(BAR(AB_SFULL), u_par);
            tc_fence_after();
            // ---- pass 1: row maximum (first maximum wins) over this warp's columns, then across the two halves
            float mx = -INFINITY; int arg = 0x7fffffff;
            for (int c0 = cbeg; c0 < cend; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(t_lane + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    if (n < p.N) {
                        float s = __uint_as_float(v[j]) * p.scale;
                        if (n < lo || n >= hi) s = ATT_MASK_VALUE;
                        if (s > mx) { mx = s; arg = n; }
                    }
                }
            }
            {   // quarters in key order and a strict comparison: the first maximum wins, like tf.argmax
                xch[half * 128 + r] = mx;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                float m4[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) m4[h] = xch[h * 128 + r];
                asm volatile("bar.sync 1, 512;" ::: "memory");
                xch[half * 128 + r] = __int_as_float(arg);
                asm volatile("bar.sync 1, 512;" ::: "memory");
                mx = m4[0]; arg = __float_as_int(xch[r]);
#pragma unroll
                for (int h = 1; h < 4; ++h) if (m4[h] > mx) { mx = m4[h]; arg = __float_as_int(xch[h * 128 + r]); }
                asm volatile("bar.sync 1, 512;" ::: "memory");
            }
            // ---- pass 2: sum of exponentials
            float sum = 0.f;
            for (int c0 = cbeg; c0 < cend; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(t_lane + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    if (n < p.N) {
                        float s = __uint_as_float(v[j]) * p.scale;
                        if (n < lo || n >= hi) s = ATT_MASK_VALUE;
                        sum += __expf(s - mx);
                    }
                }
            }
            xch[half * 128 + r] = sum;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            sum = (xch[r] + xch[128 + r]) + (xch[256 + r] + xch[384 + r]);
            const float inv = 1.f / sum;
            // ---- pass 3: probabilities -> P planes in shared memory (+ the outputs the caller asked for)
            const long long grow = (long long)b * p.T + t;
            for (int c0 = cbeg; c0 < min(p.nblk * 64, cbeg + 64); c0 += 16) {
                float a[16];
                if (c0 < p.Npad) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = c0 + j;
                        float s = __uint_as_float(v[j]) * p.scale;
                        if (n < lo || n >= hi) s = ATT_MASK_VALUE;
                        a[j] = n < p.N ? __expf(s - mx) * inv : 0.f;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[j] = 0.f;
                }
                // K-major SWIZZLE_128B: block (c0 / 64), row r, 16-byte chunks (c0 % 64) / 8 and + 1
                uint8_t* blk = arena + (c0 >> 6) * ATT_PBLK + r * 128;
                const int ch = (c0 & 63) >> 3;
                uint4 h0, l0, h1, l1;
                {
                    float t8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t8[j] = a[j];
                    split8(t8, h0, l0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) t8[j] = a[8 + j];
                    split8(t8, h1, l1);
                }
                *reinterpret_cast<uint4*>(blk + ((ch ^ (r & 7)) << 4)) = h0;
                *reinterpret_cast<uint4*>(blk + (((ch + 1) ^ (r & 7)) << 4)) = h1;
                *reinterpret_cast<uint4*>(blk + 16384 + ((ch ^ (r & 7)) << 4)) = l0;
                *reinterpret_cast<uint4*>(blk + 16384 + (((ch + 1) ^ (r & 7)) << 4)) = l1;
                if (row_ok && c0 < p.Npad) {
                    if (p.A) {
                        float* dst = p.A + grow * p.ldA + c0;
                        if (c0 + 16 <= p.N && !(p.ldA & 3)) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) if (c0 + j < p.N) dst[j] = a[j];
                        }
                    }
                    if (p.Ahi) {
                        unsigned short* dh = p.Ahi + grow * p.ldAp + c0;
                        unsigned short* dl = p.Alo + grow * p.ldAp + c0;
                        if (c0 + 16 <= p.N) {
                            *reinterpret_cast<uint4*>(dh) = h0; *reinterpret_cast<uint4*>(dh + 8) = h1;
                            *reinterpret_cast<uint4*>(dl) = l0; *reinterpret_cast<uint4*>(dl + 8) = l1;
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) if (c0 + j < p.N) st_split1(p.Ahi, p.Alo, grow * p.ldAp + c0 + j, a[j]);
                        }
                    }
                    if (p.align_t) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (c0 + j < p.N) p.align_t[((long long)b * p.N + c0 + j) * p.T + t] = a[j];
                    }
                    if (p.att_acc && t < p.maxT) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = c0 + j;
                            if (n < p.N && n < p.maxN) {
                                if (!p.G.w) att_part += a[j] * guide_w(n, t, inv_maxN, inv_maxT, inv_2g2);
                                else {
                                    const float w = guide_t(p.G, b, n, t);
                                    att_part += p.G.mse ? (a[j] - w) * (a[j] - w) : a[j] * w;
                                }
                            }
                        }
                    }
                }
            }
            if (p.argmax && row_ok && half == 0) p.argmax[grow] = arg;
            // P is complete: generic-proxy stores -> visible to the tensor core, S columns are free again
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(AB_PREADY));
            // ---- epilogue: R = P V from tensor memory -> global (fp32 + planes); this warp's quarter of the d columns
            mbar_wait(BAR(AB_RFULL), u_par);
            tc_fence_after();
            for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(t_lane + 256 + c0, v);
                tmem_ld_wait();
                if (row_ok) {
                    float* dst = p.R + grow * p.ldr + c0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    if (p.Rhi) {
                        float t8[8]; uint4 h, l;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) t8[j] = __uint_as_float(v[8 * hh + j]);
                            split8(t8, h, l);
                            *reinterpret_cast<uint4*>(p.Rhi + grow * p.ldrp + c0 + 8 * hh) = h;
                            *reinterpret_cast<uint4*>(p.Rlo + grow * p.ldrp + c0 + 8 * hh) = l;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(AB_REMPTY));
            u_par ^= 1;
        }
        if (p.att_acc) {
            att_part = warp_sum(att_part);
            if (lane == 0) atomicAdd(p.att_acc, (double)att_part);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ATT_SW + 1) tmem_dealloc1<512>(tmem_base);
}

}  // namespace oph
