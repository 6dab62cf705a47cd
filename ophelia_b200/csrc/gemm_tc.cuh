// Generic implicit-GEMM on tcgen05 with 3-term split-bf16 operands (fp32-grade accuracy):
//     C[m][n] (+)= alpha * sum_kk A[m][kk] * B[n][kk]  (+ bias[n] + addend[m][n])
//     A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi   (bf16 operands, fp32 accumulate in TMEM)
// One CTA owns a 128-row x (256*NH)-column accumulator tile in tensor memory.  fp32 activations are split
// into (hi, lo) bf16 planes by the producer warps on their way into SWIZZLE_128B shared-memory tiles;
// pre-packed weight images arrive through the bulk-copy (TMA) engine.  One thread issues tcgen05.mma.
#pragma once
#include "oph_ptx.cuh"

namespace oph {

enum { A_KMAJOR = 0, A_MNMAJOR = 1 };
enum { B_PACKED = 0, B_KMAJOR = 1, B_MNMAJOR = 2 };
enum { Z_NONE = 0, Z_BATCH = 1, Z_SPLITK = 2 };

struct OperandMap {        // logical row r -> (item b, step t) = divmod(r, L);  source step ts = t*mul + off[tap]
    const float* ptr;      // valid iff 0 <= ts < Ls;  source row = b*Ls + ts
    long long ld;          // row stride in elements (multiple of 4, rows 16-byte aligned)
    int L, Ls, mul;
    int off[3];
};

struct GemmArgs {
    int a_mode, b_mode;
    OperandMap A, Bm;
    const void* Bpacked;   // B_PACKED: [nblock][kb][half][hi 32 KiB | lo 32 KiB] shared-memory images
    int M, N;              // valid output rows / columns
    int Kc;                // reduction extent per tap (channels for conv-style A, rows for MN-major A)
    int ntaps;             // taps looped inside the CTA (conv-style A)
    int z_mode;            // grid.z meaning
    long long a_zs, b_zs, c_zs;
    int k_chunk;           // Z_SPLITK: reduction rows per z slice (multiple of 64)
    int ytaps;             // taps spread over grid.y (wgrad); output moves by c_tap_stride per tap
    long long c_tap_stride;
    float* C;
    long long ldc;
    int c_mul, c_off;      // output row = m*c_mul + c_off
    const float* bias;     // [N] or null
    const float* addend;   // [rows][ld_add] or null, indexed like C
    long long ld_add;
    float alpha;
    int atomic;            // 1: atomicAdd into C (split-K)
    int tag;               // host-side profiling category (OPH_TAG_*)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_BNH = 256;
constexpr int A_PLANE = GEMM_BM * GEMM_BK * 2;      // 16 KiB  (one bf16 plane)
constexpr int B_PLANE = GEMM_BNH * GEMM_BK * 2;     // 32 KiB
constexpr int A_SLOT = 2 * A_PLANE;                 // hi + lo
constexpr int B_SLOT = 2 * B_PLANE;
constexpr int NA_SLOTS = 3;                         // A ring depth
constexpr int NB_SLOTS = 2;                         // B ring depth
constexpr int GEMM_THREADS = 320;                   // 8 producer/epilogue warps + MMA warp + bulk-copy warp
constexpr int GEMM_SMEM = NA_SLOTS * A_SLOT + NB_SLOTS * B_SLOT + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GEMM_MAX_CLUSTER = 4;                 // CTAs (consecutive M tiles) sharing multicast weight stages

__device__ __forceinline__ void load8(const float* src, bool row_ok, int first, int limit, float (&v)[8]) {
    if (row_ok && first + 8 <= limit) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (row_ok && first + e < limit) ? __ldg(src + e) : 0.f;
    }
}

__device__ __forceinline__ void store_split(uint8_t* plane_hi, uint8_t* plane_lo, uint32_t off, const float (&v)[8]) {
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(plane_hi + off) = hi;
    *reinterpret_cast<uint4*>(plane_lo + off) = lo;
}

// map a logical reduction/output row through an OperandMap; returns false when the source row is padding
__device__ __forceinline__ bool map_row(const OperandMap& o, int r, int tap, long long& src_row) {
    int b = r / o.L;
    int t = r - b * o.L;
    int ts = t * o.mul + o.off[tap];
    src_row = (long long)b * o.Ls + ts;
    return ts >= 0 && ts < o.Ls;
}

// barrier indices
constexpr int BAR_FULL_A = 0;                          // [NA_SLOTS]
constexpr int BAR_EMPTY_A = BAR_FULL_A + NA_SLOTS;     // [NA_SLOTS]
constexpr int BAR_FULL_B = BAR_EMPTY_A + NA_SLOTS;     // [NB_SLOTS]
constexpr int BAR_EMPTY_B = BAR_FULL_B + NB_SLOTS;     // [NB_SLOTS]
constexpr int BAR_ACCUM = BAR_EMPTY_B + NB_SLOTS;
constexpr int NUM_BARS = BAR_ACCUM + 1;

template <int NH>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16x3_kernel(const __grid_constant__ GemmArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + NA_SLOTS * A_SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NA_SLOTS * A_SLOT + NB_SLOTS * B_SLOT);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NT = GEMM_BNH * NH;
    const int nblocks = (p.N + NT - 1) / NT;
    const int nb = blockIdx.y % nblocks;
    const int ytap = blockIdx.y / nblocks;
    const int m0 = blockIdx.x * GEMM_BM;               // may lie beyond M for cluster-padding CTAs (they only feed the ring)
    const int n0 = nb * NT;
    const int z = blockIdx.z;
    const uint32_t csize = cluster_nctarank();         // CTAs of one cluster share the packed-B stages by multicast
    const uint32_t crank = cluster_ctarank();
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);

    int k_begin = 0, k_end = p.Kc;
    long long a_z = 0, b_z = 0, c_z = 0;
    if (p.z_mode == Z_BATCH) { a_z = z * p.a_zs; b_z = z * p.b_zs; c_z = z * p.c_zs; }
    if (p.z_mode == Z_SPLITK) { k_begin = z * p.k_chunk; k_end = min(p.Kc, k_begin + p.k_chunk); }
    const int KBc = (k_end - k_begin + GEMM_BK - 1) / GEMM_BK;
    const int ntl = (p.a_mode == A_KMAJOR) ? p.ntaps : 1;
    const int KB = ntl * KBc;
    if (KB <= 0) return;                               // uniform over the cluster (depends on z only)

    if (tid == 0) {
        const uint32_t nprod = 8;
        for (int i = 0; i < NA_SLOTS; ++i) { mbar_init(BAR(BAR_FULL_A + i), nprod); mbar_init(BAR(BAR_EMPTY_A + i), 1); }
        const uint32_t full_b = (p.b_mode == B_PACKED) ? 1u : nprod;
        const uint32_t empty_b = (p.b_mode == B_PACKED) ? csize : 1u;   // every CTA of the cluster must release a multicast slot
        for (int i = 0; i < NB_SLOTS; ++i) { mbar_init(BAR(BAR_FULL_B + i), full_b); mbar_init(BAR(BAR_EMPTY_B + i), empty_b); }
        mbar_init(BAR(BAR_ACCUM), 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 8) tmem_alloc<NT>(smem_u32(tmem_slot));
    tc_fence_before();
    if (csize > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ================================================================ producers
        const float* Ap = p.A.ptr + a_z;
        const float* Bp = (p.b_mode == B_PACKED) ? nullptr : (p.Bm.ptr + b_z);
        int a_t[4]; long long a_base[4]; bool a_ok[4];   // conv-style A: 4 rows per thread, fixed for the whole tile
        if (p.a_mode == A_KMAJOR) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int g = m0 + (tid >> 3) + 32 * i;
                a_ok[i] = g < p.M;
                int b = g / p.A.L;
                a_t[i] = g - b * p.A.L;
                a_base[i] = (long long)b * p.A.Ls;
            }
        }
        // global -> register load of this thread's 4 chunks of the A tile of k-block kb
        auto load_a = [&](int kb, float (&v)[4][8]) {
            const int tap = (p.a_mode == A_KMAJOR) ? kb / KBc : ytap;
            const int kk0 = k_begin + (kb - (kb / KBc) * KBc) * GEMM_BK;
            if (p.a_mode == A_KMAJOR) {
                const int c = kk0 + (tid & 7) * 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ts = a_t[i] * p.A.mul + p.A.off[tap];
                    const bool ok = a_ok[i] && ts >= 0 && ts < p.A.Ls;
                    load8(Ap + (a_base[i] + ts) * p.A.ld + c, ok, c, k_end, v[i]);
                }
            } else {
                const int m = m0 + (tid & 15) * 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = kk0 + (tid >> 4) + 16 * i;
                    long long srow;
                    const bool ok = map_row(p.A, r, tap, srow) && r < k_end;
                    load8(Ap + srow * p.A.ld + m, ok, m, p.M, v[i]);
                }
            }
        };
        uint32_t a_soff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (p.a_mode == A_KMAJOR) {
                const int rl = (tid >> 3) + 32 * i, chunk = tid & 7;
                a_soff[i] = rl * 128 + ((chunk ^ (rl & 7)) << 4);
            } else {
                const int rl = (tid >> 4) + 16 * i, j = tid & 15;
                a_soff[i] = (j >> 3) * 8192 + (rl >> 3) * 1024 + (rl & 7) * 128 + (((j & 7) ^ (rl & 7)) << 4);
            }
        }
        float va[4][8], vn[4][8];
        load_a(0, va);
        for (int kb = 0; kb < KB; ++kb) {
            if (kb + 1 < KB) load_a(kb + 1, vn);           // next tile's loads are in flight while this one is converted
            {
                const int slot = kb % NA_SLOTS;
                mbar_wait(BAR(BAR_EMPTY_A + slot), ((kb / NA_SLOTS) & 1) ^ 1);
                uint8_t* hi = sA + slot * A_SLOT;
#pragma unroll
                for (int i = 0; i < 4; ++i) store_split(hi, hi + A_PLANE, a_soff[i], va[i]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(BAR_FULL_A + slot));
            }
            // ---------------- B tile(s) from fp32 activations
            if (p.b_mode != B_PACKED) {
                const int tap = (p.a_mode == A_KMAJOR) ? kb / KBc : ytap;
                const int kk0 = k_begin + (kb - (kb / KBc) * KBc) * GEMM_BK;
#pragma unroll 1
                for (int h = 0; h < NH; ++h) {
                    const int bi = kb * NH + h, slot = bi % NB_SLOTS;
                    uint8_t* hi = sB + slot * B_SLOT;
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {      // 2 x 4 chunks per thread keeps registers bounded
                        float v[4][8];
                        uint32_t soff[4];
                        if (p.b_mode == B_KMAJOR) {
                            const int chunk = tid & 7, c = kk0 + chunk * 8;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int nl = (tid >> 3) + 32 * (i + 4 * half);
                                const int n = n0 + h * GEMM_BNH + nl;
                                load8(Bp + (long long)n * p.Bm.ld + c, n < p.N, c, k_end, v[i]);
                                soff[i] = nl * 128 + ((chunk ^ (nl & 7)) << 4);
                            }
                        } else {
                            const int j = tid & 31, n = n0 + h * GEMM_BNH + j * 8;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int rl = (tid >> 5) + 8 * (i + 4 * half);
                                const int r = kk0 + rl;
                                long long srow;
                                const bool ok = map_row(p.Bm, r, tap, srow) && r < k_end;
                                load8(Bp + srow * p.Bm.ld + n, ok, n, p.N, v[i]);
                                soff[i] = (j >> 3) * 8192 + (rl >> 3) * 1024 + (rl & 7) * 128 + (((j & 7) ^ (rl & 7)) << 4);
                            }
                        }
                        if (half == 0) mbar_wait(BAR(BAR_EMPTY_B + slot), ((bi / NB_SLOTS) & 1) ^ 1);
#pragma unroll
                        for (int i = 0; i < 4; ++i) store_split(hi, hi + B_PLANE, soff[i], v[i]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(BAR_FULL_B + slot));
                }
            }
            if (kb + 1 < KB) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < 8; ++e) va[i][e] = vn[i][e];
            }
        }
    } else if (warp == 8) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(GEMM_BM, GEMM_BNH, p.a_mode == A_MNMAJOR, p.b_mode == B_MNMAJOR);
            const uint32_t a_step = (p.a_mode == A_MNMAJOR) ? 2048u : 32u;
            const uint32_t a_lbo = (p.a_mode == A_MNMAJOR) ? 8192u : 16u;
            const uint32_t b_step = (p.b_mode == B_MNMAJOR) ? 2048u : 32u;
            const uint32_t b_lbo = (p.b_mode == B_MNMAJOR) ? 8192u : 16u;
            const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
            const bool mcast = (p.b_mode == B_PACKED) && csize > 1;
            for (int kb = 0; kb < KB; ++kb) {
                const int as = kb % NA_SLOTS;
                mbar_wait(BAR(BAR_FULL_A + as), (kb / NA_SLOTS) & 1);
                const uint32_t a_hi = sA_addr + as * A_SLOT, a_lo = a_hi + A_PLANE;
                for (int h = 0; h < NH; ++h) {
                    const int bi = kb * NH + h, bs = bi % NB_SLOTS;
                    mbar_wait(BAR(BAR_FULL_B + bs), (bi / NB_SLOTS) & 1);
                    tc_fence_after();
                    const uint32_t b_hi = sB_addr + bs * B_SLOT, b_lo = b_hi + B_PLANE;
                    const uint32_t d = tmem_base + h * GEMM_BNH;
#pragma unroll
                    for (int ks = 0; ks < GEMM_BK / 16; ++ks) {
                        const uint64_t dah = make_sdesc(a_hi + ks * a_step, a_lbo, 1024);
                        const uint64_t dal = make_sdesc(a_lo + ks * a_step, a_lbo, 1024);
                        const uint64_t dbh = make_sdesc(b_hi + ks * b_step, b_lbo, 1024);
                        const uint64_t dbl = make_sdesc(b_lo + ks * b_step, b_lbo, 1024);
                        umma_bf16(d, dah, dbh, idesc, (kb | ks) != 0);
                        umma_bf16(d, dah, dbl, idesc, 1);
                        umma_bf16(d, dal, dbh, idesc, 1);
                    }
                    // B slot free once these MMAs retire; a multicast slot must be released in every CTA of the cluster
                    if (mcast) umma_commit_mcast(BAR(BAR_EMPTY_B + bs), cmask); else umma_commit(BAR(BAR_EMPTY_B + bs));
                }
                umma_commit(BAR(BAR_EMPTY_A + as));    // A slot free
            }
            umma_commit(BAR(BAR_ACCUM));               // accumulator complete
        }
        __syncwarp();
    } else {
        // ================================================================ packed-weight loader (bulk copy engine)
        if (lane == 0 && p.b_mode == B_PACKED) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.Bpacked) + (size_t)nb * KB * NH * B_SLOT;
            const int total = KB * NH;
            const uint32_t share = B_SLOT / csize;      // each CTA fetches 1/csize of a stage and multicasts it
            for (int bi = 0; bi < total; ++bi) {
                const int slot = bi % NB_SLOTS;
                mbar_wait(BAR(BAR_EMPTY_B + slot), ((bi / NB_SLOTS) & 1) ^ 1);
                mbar_arrive_expect_tx(BAR(BAR_FULL_B + slot), B_SLOT);
                const uint32_t dst = smem_u32(sB + slot * B_SLOT) + crank * share;
                const uint8_t* s_ = src + (size_t)bi * B_SLOT + crank * share;
                if (csize > 1) bulk_g2s_mcast(dst, s_, share, BAR(BAR_FULL_B + slot), cmask);
                else bulk_g2s(dst, s_, share, BAR(BAR_FULL_B + slot));
            }
        }
        __syncwarp();
    }

    // ==================================================================== epilogue (warps 0..7)
    if (warp < 8) {
        mbar_wait(BAR(BAR_ACCUM), 0);
        tc_fence_after();
        const int q = warp & 3, colhalf = warp >> 2;
        float* stage = reinterpret_cast<float*>(sA) + warp * (32 * 33);   // operand ring is idle now
        float* Cb = p.C + c_z + (long long)ytap * p.c_tap_stride;
        const float* addb = p.addend ? p.addend + c_z + (long long)ytap * p.c_tap_stride : nullptr;
        constexpr int CH = NT / 2 / 32;
        if (m0 + q * 32 < p.M) {
            for (int ch = 0; ch < CH; ++ch) {
                const int col0 = colhalf * (NT / 2) + ch * 32;
                if (n0 + col0 >= p.N) break;           // warp-uniform
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + col0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                const int gcol = n0 + col0 + lane;
                const bool col_ok = gcol < p.N;
                const float bv = (p.bias && col_ok) ? __ldg(p.bias + gcol) : 0.f;
                for (int rr = 0; rr < 32; ++rr) {
                    const int grow = m0 + q * 32 + rr;
                    if (grow >= p.M) break;            // warp-uniform
                    if (col_ok) {
                        const long long crow = (long long)grow * p.c_mul + p.c_off;
                        float val = stage[rr * 33 + lane] * p.alpha + bv;
                        if (addb) val += __ldg(addb + crow * p.ld_add + gcol);
                        float* dst = Cb + crow * p.ldc + gcol;
                        if (p.atomic) atomicAdd(dst, val); else *dst = val;
                    }
                }
                __syncwarp();
            }
        }
        tc_fence_before();
    }
    // peers may still multicast-arrive on this CTA's barriers until they are done too
    if (csize > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 8) tmem_dealloc<NT>(tmem_base);
}

}  // namespace oph
