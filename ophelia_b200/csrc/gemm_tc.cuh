// Generic implicit-GEMM on tcgen05 with 3-term split-bf16 operands (fp32-grade accuracy):
//     C[m][n] (+)= alpha * sum_kk A[m][kk] * B[n][kk]  (+ bias[n] + addend[m][n])
//     A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi   (bf16 operands, fp32 accumulate in TMEM)
//
// PERSISTENT kernel, one CTA pair (cluster of 2, tcgen05 cta_group::2) per two SMs.  A work unit is a 256-row x
// 256-column output tile (x one tap / batch item / split-K slice): each CTA of the pair owns 128 rows of the fp32
// accumulator in its tensor memory and stages its own 128 rows of A plus HALF of every B stage; the MMA unit reads
// the other half from the partner's shared memory.  The 512 TMEM columns hold TWO accumulator stages, so the
// epilogue of unit i overlaps the main loop of unit i+1.  Operands normally arrive as TMA tensor copies of
// split-bf16 planes (activations: written by the row-wise kernels; weights: pre-packed smem images) that signal the
// leader CTA's barriers directly; fp32 operands without planes are split by 16 producer warps on their way into the
// SWIZZLE_128B tiles.  One thread of the leader CTA issues every tcgen05.mma; when nothing needs producing, the
// producer warps drain the accumulators (16-warp epilogue).
#pragma once
#include <cuda.h>
#include "oph_ptx.cuh"

namespace oph {

enum { A_KMAJOR = 0, A_MNMAJOR = 1 };
enum { B_PACKED = 0, B_KMAJOR = 1, B_MNMAJOR = 2 };
enum { Z_NONE = 0, Z_BATCH = 1, Z_SPLITK = 2 };

// Division by a launch constant as multiply-high + shift (valid for 0 <= n < 2^31): the unit decode of every role sits on
// the start-up path of a launch, and the emulated integer division is ~100+ dependent cycles a piece.
struct FastDiv { uint32_t mul, shr; };     // mul == 0: divisor 1
inline FastDiv make_fastdiv(int d) {
    FastDiv f{0u, 0u};
    if (d <= 1) return f;
    int lg = 0;
    while ((1ll << lg) < d) ++lg;                      // ceil(log2 d)
    const unsigned p = 31u + (unsigned)lg;
    f.mul = (uint32_t)(((1ull << p) + (unsigned)d - 1ull) / (unsigned)d);
    f.shr = p - 32u;
    return f;
}
__device__ __forceinline__ int fdiv(int n, FastDiv f) { return f.mul ? (int)(__umulhi((uint32_t)n, f.mul) >> f.shr) : n; }

struct OperandMap {        // logical row r -> (item b, step t) = divmod(r, L);  source step ts = t*mul + off[tap]
    const float* ptr;      // valid iff 0 <= ts < Ls;  source row = b*Ls + ts
    const unsigned short* hi;   // optional pre-split operand: bf16 planes hi / lo with the same [row][ld] indexing
    const unsigned short* lo;   //   (written by the row-wise kernels); then ptr is unused and ld % 8 == 0
    long long ld;          // row stride in elements (fp32: multiple of 4; planes: multiple of 8)
    int L, Ls, mul;
    int off[3];
};

struct GemmArgs {
    int a_mode, b_mode;
    OperandMap A, Bm;
    const void* Bpacked;   // B_PACKED: [n-block of 256][kb][cta 0: hi 16 KiB | lo 16 KiB][cta 1: hi | lo] smem images
    int M, N;              // valid output rows / columns
    int Kc;                // reduction extent per tap (channels for conv-style A, rows for MN-major A)
    int ntaps;             // taps looped inside the CTA (conv-style A)
    int z_mode;            // meaning of the z index of a work unit
    int zdim;              // number of z slices (batch items or split-K slices)
    long long a_zs, b_zs, c_zs;
    int k_chunk;           // Z_SPLITK: reduction rows per z slice (multiple of 64)
    int ytaps;             // taps spread over grid.y (wgrad); output moves by c_tap_stride per tap
    long long c_tap_stride;
    float* C;
    long long ldc;
    unsigned short* Chi;   // optional: the output also as split-bf16 planes (operand format of the next product);
    unsigned short* Clo;   //   plain (non-atomic) 16-byte-aligned outputs only
    long long ldcp, cp_zs; // plane row stride / z stride in elements
    int c_mul, c_off;      // output row = m*c_mul + c_off
    const float* bias;     // [N] or null
    const float* addend;   // [rows][ld_add] or null, indexed like C
    long long ld_add;
    float alpha;
    int atomic;            // 1: atomicAdd into C (split-K)
    int tag;               // host-side profiling category (OPH_TAG_*)
    long long* dbg;        // optional [pairs][8] cycle counters of the MMA / producer waits (diagnostics)
    int prof_k;            // host-side: true reduction length for the FLOP count when Kc is in k-blocks (r_tma)
    int dbg_flags;         // diagnostics only: 1 = producers skip data movement, 2 = weight loader skips copies
    // TMA feed of a pre-split conv-style A operand: FLAT 128-row tiles over the (channels, items*time) tensor map (the
    // fewest tiles, no per-item tail tiles).  A tap shifts the row coordinate; rows shifted beyond either end of the
    // whole tensor are out-of-range zeros, rows shifted across an item boundary are zeroed in shared memory by the
    // fix-up warp before the MMA sees the stage.
    // Weight-gradient products with both operands pre-split (r_tma): the reduction runs over (item, 64-step block) pairs
    // so that tap shifts and item boundaries are out-of-range coordinates too; k_chunk then counts k-blocks.
    int a_tma;             // K-major A tiles come from tmA_hi / tmA_lo: 1 = flat conv-style tiles (2-D map), 2 = tiles of
                           //   batch item z (3-D map, Z_BATCH products such as the attention contractions)
    int b_tma;             // activation B operand of item z from tmB_hi / tmB_lo: 2 = K-major rows, 3 = MN-major steps
    int r_tma;             // 1: MN-major A and B tiles (64 steps x 128 channels each) come from tmA_* / tmB_*; the
                           //   reduction runs over the steps of item z (Z_BATCH) or over (item, block) pairs (Z_SPLITK)
    int items;             // number of batch items (M = items * A.L)
    // remainder K-split (RED outputs only): work items >= split_from are split_s k-block slices of the remaining units,
    // so that the last round of the persistent CTA pairs is short instead of mostly idle
    int split_from, split_s;
    size_t zero_bytes;     // host-side: > 0 if launch_gemm may memset C[0, zero_bytes) (the caller owns a dense output)
    // conv tail fused into the epilogue (outputs with at most 256 columns: one work unit holds whole rows):
    // y = dropout(act(LN(z))), z = acc + bias; C (z, may be NULL at inference), ln_stats, ln_y, planes, ln_ysig
    int ln_fuse;           // set by launch_gemm when the request (ln_gamma != NULL) can be honoured
    const float* ln_gamma; const float* ln_beta;
    float* ln_stats;       // [M][2] (mean, rstd) or NULL
    float* ln_y; long long ln_ldy;
    unsigned short* ln_yhi; unsigned short* ln_ylo; long long ln_ldp;
    float* ln_ysig; long long ln_ldys;
    int ln_act; float ln_drop; unsigned long long ln_seed; const long long* ln_step;
    int split_red;         // 1: C was zeroed by the host and only the split work items accumulate with RED (plain outputs)
    // Highway tail fused into the same launch (modules.hc, N = 2C): the first hc_fused row-tile pairs (256 rows each) are
    // SUPER-UNITS -- a CTA pair runs all column blocks of such a tile back to back, its epilogue warps write z as usual and,
    // after the last block, finish their CTA's 128 rows from the just-written z (L2 hits): LN(H1), LN(H2), sigmoid gate,
    // highway mix with x, dropout -> y, operand planes, row statistics.  Row tiles >= hc_fused stay plain units (the host
    // launches the stand-alone tail kernel for their rows).
    int hc_fused, hc_C;
    const float* hc_x; long long hc_ldx;
    const float* hc_g1; const float* hc_b1; const float* hc_g2; const float* hc_b2;
    float* hc_y; long long hc_ldy;
    unsigned short* hc_yhi; unsigned short* hc_ylo; long long hc_ldp;
    float* hc_stats;       // [M][4] (mean1, rstd1, mean2, rstd2) or NULL
    float hc_drop; unsigned long long hc_seed; const long long* hc_step;
    // launch constants as fast divisors: row-tile pairs, column blocks, grid taps, slices per remainder unit, steps per item
    // (A.L), k-blocks per tap, k-blocks per unit, 64-step blocks per item
    FastDiv fd_MP, fd_nb, fd_yt, fd_ss, fd_L, fd_KBc, fd_KB, fd_KBI;
    alignas(64) CUtensorMap tmA_hi;
    alignas(64) CUtensorMap tmA_lo;
    alignas(64) CUtensorMap tmB_hi;
    alignas(64) CUtensorMap tmB_lo;
};

constexpr int GEMM_BM = 128;                        // rows per CTA (256 per pair)
constexpr int GEMM_BK = 64;
constexpr int GEMM_BN = 256;                        // columns per work unit = one MMA instruction's N
constexpr int GEMM_BNC = GEMM_BN / 2;               // B rows staged by each CTA of the pair
constexpr int A_PLANE = GEMM_BM * GEMM_BK * 2;      // 16 KiB  (one bf16 plane)
constexpr int B_PLANE = GEMM_BNC * GEMM_BK * 2;     // 16 KiB
constexpr int A_SLOT = 2 * A_PLANE;                 // hi + lo
constexpr int B_SLOT = 2 * B_PLANE;                 // per CTA
constexpr int B_STAGE = 2 * B_SLOT;                 // both CTAs: one k-block of the packed image of a 256-column block
constexpr int NA_SLOTS = 3;
constexpr int NB_SLOTS = 3;
constexpr int N_ACC = 2;                            // accumulator stages in tensor memory (2 x 256 columns)
constexpr int NPW = 16;                             // producer warps
constexpr int GEMM_THREADS = (NPW + 8) * 32;        // 16 producer (or epilogue) warps, 4 epilogue warps, then MMA issuer / copy loader / fix-up / idle
constexpr int EPI_WARPS = 16;                       // epilogue workers when both operands come from the copy engines
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 2048;   // per-warp [32 rows][16 columns] fp32 transposition buffers
constexpr int GEMM_SMEM = NA_SLOTS * A_SLOT + NB_SLOTS * B_SLOT + EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GEMM_MAX_PAIRS = 74;
constexpr int DBG_STRIDE = 24;                       // long long slots per CTA pair in the diagnostics buffer                  // 148 SMs

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

// one 8-element chunk of an operand into registers: fp32 source -> 8 floats; pre-split source -> hi uint4 | lo uint4
__device__ __forceinline__ void load_chunk(const OperandMap& o, long long z_off, long long off, bool ok, int first, int limit, float (&v)[8]) {
    if (o.hi) {
        uint4 h = make_uint4(0u, 0u, 0u, 0u), l = h;
        if (ok && first < limit) {
            h = __ldg(reinterpret_cast<const uint4*>(o.hi + z_off + off));
            l = __ldg(reinterpret_cast<const uint4*>(o.lo + z_off + off));
        }
        v[0] = __uint_as_float(h.x); v[1] = __uint_as_float(h.y); v[2] = __uint_as_float(h.z); v[3] = __uint_as_float(h.w);
        v[4] = __uint_as_float(l.x); v[5] = __uint_as_float(l.y); v[6] = __uint_as_float(l.z); v[7] = __uint_as_float(l.w);
    } else {
        const float* src = o.ptr + z_off + off;
        if (ok && first + 8 <= limit) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (ok && first + e < limit) ? __ldg(src + e) : 0.f;
        }
    }
}
// registers -> the hi / lo planes of a shared-memory tile (conversion only for fp32 sources)
__device__ __forceinline__ void store_chunk(bool split_src, uint8_t* plane_hi, uint8_t* plane_lo, uint32_t off, const float (&v)[8]) {
    if (split_src) {
        *reinterpret_cast<uint4*>(plane_hi + off) = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
        *reinterpret_cast<uint4*>(plane_lo + off) = make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
    } else {
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(plane_hi + off) = hi;
        *reinterpret_cast<uint4*>(plane_lo + off) = lo;
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

// barrier indices (same layout in both CTAs; FULL_* / T_EMPTY are only used in the leader, LAND_A in either CTA)
constexpr int BAR_FULL_A = 0;                          // [NA_SLOTS] leader: 2*NPW producer-warp arrivals, or 1 + copy bytes of both CTAs
constexpr int BAR_EMPTY_A = BAR_FULL_A + NA_SLOTS;     // [NA_SLOTS] leader's commit, multicast to both CTAs
constexpr int BAR_FULL_B = BAR_EMPTY_A + NA_SLOTS;     // [NB_SLOTS]
constexpr int BAR_EMPTY_B = BAR_FULL_B + NB_SLOTS;     // [NB_SLOTS]
constexpr int BAR_LAND_A = BAR_EMPTY_B + NB_SLOTS;     // [NA_SLOTS] local: an A tile that needs the item-boundary fix-up landed
constexpr int BAR_T_FULL = BAR_LAND_A + NA_SLOTS;      // [N_ACC] accumulator stage complete (commit, multicast)
constexpr int BAR_T_EMPTY = BAR_T_FULL + N_ACC;        // [N_ACC] accumulator stage drained by 4+4 epilogue warps
constexpr int NUM_BARS = BAR_T_EMPTY + N_ACC;

struct Unit {               // one 256x256 output tile of one tap / z slice
    int m0, n0, nb, ytap, k_begin, k_end, KBc, KB;
    int kb0, kb1;           // k-block range of this work item (the whole unit unless it is a remainder slice)
    int z;                  // batch item (Z_BATCH) or split-K slice
    int sliced;             // 1: this work item is one k-block slice of a remainder unit
    int rows;               // valid rows of this CTA's 128-row tile (<= 0: padding CTA)
    long long a_z, b_z, c_z;
};

// does the 128-row flat tile starting at row m0 contain rows whose tap-shifted step leaves their batch item?
__device__ __forceinline__ bool needs_fix(int m0, int off, int L, FastDiv fdL) {
    if (off == 0) return false;
    const int t0 = m0 - fdiv(m0, fdL) * L;
    if (t0 + GEMM_BM > L) return true;                 // the tile crosses an item boundary
    return off < 0 ? t0 < -off : t0 + GEMM_BM > L - off;
}

__device__ __forceinline__ Unit decode_unit(const GemmArgs& p, int v, int MP, int nblocks, uint32_t crank) {
    Unit t;
    int u = v, slice = 0, nslices = 1;
    t.sliced = 0;
    if (p.split_s > 1 && v >= p.split_from) {
        t.sliced = 1;
        const int w = v - p.split_from, wq = fdiv(w, p.fd_ss);
        u = p.split_from + wq; slice = w - wq * p.split_s; nslices = p.split_s;
    }
    int rest = fdiv(u, p.fd_MP);
    const int mp = u - rest * MP;
    const int r1 = fdiv(rest, p.fd_nb);
    t.nb = rest - r1 * nblocks;
    const int z = fdiv(r1, p.fd_yt);
    t.ytap = r1 - z * p.ytaps;
    t.z = z;
    const int mt = mp * 2 + (int)crank;
    t.m0 = mt * GEMM_BM;                               // may lie beyond M for the padding CTA of the last pair
    t.rows = min(GEMM_BM, p.M - t.m0);
    t.n0 = t.nb * GEMM_BN;
    t.k_begin = 0; t.k_end = p.Kc; t.a_z = t.b_z = t.c_z = 0;
    if (p.z_mode == Z_BATCH) { t.a_z = z * p.a_zs; t.b_z = z * p.b_zs; t.c_z = z * p.c_zs; }
    if (p.z_mode == Z_SPLITK) { t.k_begin = z * p.k_chunk; t.k_end = min(p.Kc, t.k_begin + p.k_chunk); }
    if (p.r_tma) {                                     // k_begin / k_end count (item, 64-step block) pairs
        t.KBc = t.KB = t.k_end - t.k_begin;
    } else {
        t.KBc = (t.k_end - t.k_begin + GEMM_BK - 1) / GEMM_BK;
        t.KB = ((p.a_mode == A_KMAJOR) ? p.ntaps : 1) * t.KBc;
    }
    if (nslices == 1) { t.kb0 = 0; t.kb1 = t.KB; }
    else { t.kb0 = fdiv(slice * t.KB, p.fd_ss); t.kb1 = fdiv((slice + 1) * t.KB, p.fd_ss); }
    return t;
}

// The it-th work item of CTA pair `pair` as a unit index for decode_unit, or -1 past the end.  Plain launches walk the
// units round-robin.  With a fused highway tail the first hc_fused row-tile pairs are super-units (all column blocks of a
// row tile consecutively on one pair), the other row tiles follow as plain units.
template <bool HC>
__device__ __forceinline__ int item_unit(const GemmArgs& p, int it, int pair, int npairs, int MP, int nblocks, int total) {
    if (!HC || !p.hc_fused) { const int u = pair + it * npairs; return u < total ? u : -1; }
    const int F = p.hc_fused;
    const int nsup = F > pair ? (F - pair + npairs - 1) / npairs : 0;
    if (it < nsup * nblocks) { const int q = fdiv(it, p.fd_nb); return (it - q * nblocks) * MP + pair + q * npairs; }
    const int r = pair + (it - nsup * nblocks) * npairs, W = MP - F;
    if (W <= 0 || r >= W * nblocks) return -1;
    return (r / W) * MP + F + r % W;
}
// is item `it` of this pair the last column block of a fused super-unit?
template <bool HC>
__device__ __forceinline__ bool item_closes_tile(const GemmArgs& p, int it, int pair, int npairs, int nblocks) {
    if (!HC || !p.hc_fused) return false;
    const int nsup = p.hc_fused > pair ? (p.hc_fused - pair + npairs - 1) / npairs : 0;
    return it < nsup * nblocks && (it - fdiv(it, p.fd_nb) * nblocks) == nblocks - 1;
}

__device__ __forceinline__ float4 ld_cg4(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// Highway tail over `nrows` rows of one CTA (the 16 epilogue warps, 512 threads), same arithmetic in the same order as
// hc_post_fwd_wide_kernel: WPR warps share a row (256 channels per warp, two float4 per lane), per-warp two-pass moments
// merged exactly across the warps of a row.  z rows are read back with ld.global.cg (they were written by other warps of
// this CTA just before the CTA-wide barrier); sm = the epilogue staging area, free during this phase.
template <int WPR>
__device__ __forceinline__ void hc_tail_rows(const GemmArgs& p, float* sm, int warp, int lane, long long row0, int nrows,
                                             const float* zrow0) {
    constexpr int C = 256 * WPR;
    constexpr int GROUPS = 16 / WPR;
    float* spar = sm;                                   // [4][C]: g1, b1, g2, b2
    float* sx = sm + 4 * C;                             // [2 parities][16 warps][4] partial moments
    for (int i = warp * 32 + lane; i < C; i += 512) {
        spar[i] = __ldg(p.hc_g1 + i); spar[C + i] = __ldg(p.hc_b1 + i);
        spar[2 * C + i] = __ldg(p.hc_g2 + i); spar[3 * C + i] = __ldg(p.hc_b2 + i);
    }
    asm volatile("bar.sync 13, 512;" ::: "memory");
    const int grp = warp / WPR, part = warp % WPR;
    const int cbase = part * 256 + lane * 4;
    const float inv_keep = p.hc_drop > 0.f ? 1.f / (1.f - p.hc_drop) : 1.f;
    const unsigned long long sd = eff_seed(p.hc_seed, p.hc_step);
    int it = 0;
    for (int r = grp; r < nrows; r += GROUPS, ++it) {
        const long long row = row0 + r;
        const float* zr = zrow0 + (long long)r * p.ldc;
        float4 z1[2], z2[2], xv[2], o[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            z1[i] = ld_cg4(zr + cbase + 128 * i);
            z2[i] = ld_cg4(zr + C + cbase + 128 * i);
            xv[i] = __ldg(reinterpret_cast<const float4*>(p.hc_x + row * p.hc_ldx + cbase + 128 * i));
        }
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
        if (WPR > 1) {
            float* my = sx + ((it & 1) * 16 + warp) * 4;
            if (lane == 0) *reinterpret_cast<float4*>(my) = make_float4(m1, q1, m2, q2);
            asm volatile("bar.sync %0, %1;" ::"r"(5 + grp), "r"(WPR * 32) : "memory");
            float mm1 = 0.f, mm2 = 0.f;
            float4 ow[WPR];
#pragma unroll
            for (int w = 0; w < WPR; ++w) { ow[w] = *reinterpret_cast<const float4*>(sx + ((it & 1) * 16 + grp * WPR + w) * 4); mm1 += ow[w].x; mm2 += ow[w].z; }
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
        if (p.hc_stats && lane == 0 && part == 0) *reinterpret_cast<float4*>(p.hc_stats + row * 4) = make_float4(m1, r1, m2, r2);
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
                float v = g * u2 + (1.f - g) * OPH_F4(xv[i], e);
                if (p.hc_drop > 0.f) v *= drop_scale(sd, (unsigned long long)row * C + cbase + 128 * i + e, p.hc_drop, inv_keep);
                OPH_F4(o[i], e) = v;
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            *reinterpret_cast<float4*>(p.hc_y + row * p.hc_ldy + cbase + 128 * i) = o[i];
            if (p.hc_yhi) {
                uint2 hh, ll;
                split4(o[i], hh, ll);
                *reinterpret_cast<uint2*>(p.hc_yhi + row * p.hc_ldp + cbase + 128 * i) = hh;
                *reinterpret_cast<uint2*>(p.hc_ylo + row * p.hc_ldp + cbase + 128 * i) = ll;
            }
        }
    }
    asm volatile("bar.sync 13, 512;" ::: "memory");        // the staging area goes back to the plain epilogue
}

// HC = true: the instantiation that can run the highway tail (hc_fused); the plain one carries none of that code
template <bool HC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ GemmArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + NA_SLOTS * A_SLOT;
    float* sStage = reinterpret_cast<float*>(smem + NA_SLOTS * A_SLOT + NB_SLOTS * B_SLOT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NA_SLOTS * A_SLOT + NB_SLOTS * B_SLOT + EPI_STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned long long gt_start = 0;
    if (p.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_start));
    const uint32_t crank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = partner
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int MP = ((p.M + GEMM_BM - 1) / GEMM_BM + 1) / 2;
    const int nblocks = (p.N + GEMM_BN - 1) / GEMM_BN;
    const int units = MP * nblocks * p.ytaps * p.zdim;
    const int total = p.split_s > 1 ? p.split_from + (units - p.split_from) * p.split_s : units;   // work items
    const bool packed = p.b_mode == B_PACKED;
    const bool a_k = p.a_mode == A_KMAJOR;
    // both operands arrive through the copy engines: the 16 producer warps have nothing to stage and drain the
    // accumulators instead (4 x more epilogue warps, the dedicated epilogue warps then idle)
    const bool copy_fed = (p.a_tma && (packed || p.b_tma)) || p.r_tma;
    const bool rotate = p.a_tma == 1 && packed && !(p.dbg_flags & 4);
    const bool nofix = (p.dbg_flags & (32 | 128)) != 0;        // diagnostics: skip the item-boundary fix-up (wrong results)

    if (warp == NPW + 5 && lane == 0) {                // the copy-engine loader's descriptors: fetched while the prologue runs
        if (p.a_tma || p.r_tma) { tma_prefetch_desc(&p.tmA_hi); tma_prefetch_desc(&p.tmA_lo); }
        if (p.r_tma || packed || p.b_tma) tma_prefetch_desc(&p.tmB_hi);
        if (p.r_tma || p.b_tma) tma_prefetch_desc(&p.tmB_lo);
    }
    if (tid == 0) {
        for (int i = 0; i < NA_SLOTS; ++i) {
            // copies: the leader's expect_tx arrive (+ for conv-style tiles one token per CTA: sent by its loader when
            // the tile needs no fix-up, by its fix-up warp otherwise)
            mbar_init(BAR(BAR_FULL_A + i), p.a_tma ? 3 : (p.r_tma ? 1 : 2 * NPW));
            mbar_init(BAR(BAR_EMPTY_A + i), 1);
            mbar_init(BAR(BAR_LAND_A + i), 1);
        }
        for (int i = 0; i < NB_SLOTS; ++i) {
            mbar_init(BAR(BAR_FULL_B + i), (packed || p.r_tma || p.b_tma) ? 1 : 2 * NPW);
            mbar_init(BAR(BAR_EMPTY_B + i), 1);
        }
        for (int i = 0; i < N_ACC; ++i) { mbar_init(BAR(BAR_T_FULL + i), 1); mbar_init(BAR(BAR_T_EMPTY + i), copy_fed ? 2 * EPI_WARPS : 8); }
        mbar_fence_init();
        fence_proxy_async();
        if (p.dbg && crank == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); p.dbg[pair * DBG_STRIDE + 17] = (long long)(gt_ - gt_start); }
    }
    if (warp == NPW + 4) {
        tmem_alloc2<N_ACC * GEMM_BN>(smem_u32(tmem_slot));
        if (p.dbg && crank == 0 && lane == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); p.dbg[pair * DBG_STRIDE + 16] = (long long)(gt_ - gt_start); }
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the prologue above overlapped the predecessor kernel's tail; its results are visible from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p.dbg && crank == 0 && tid == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[pair * DBG_STRIDE + 9] = (long long)(gt - gt_start);      // ns from kernel entry to the end of the prologue
    }


    // ================================================================ epilogue worker: TMEM -> global
    // One warp drains the 32 accumulator rows of its TMEM lane quadrant q for the 16-column blocks cg, cg + ncg, ...
    // Each block goes through a swizzled [32][16] shared-memory buffer so that a lane ends up with 4 consecutive
    // columns of one row: every global access is a 16-byte vector and a warp instruction covers 8 rows x 64 bytes.
    auto epilogue = [&](const int q, const int cg, const int ncg, float* stage) {
        const int cb_last = cg + ncg * ((GEMM_BN / 16 - 1 - cg) / ncg);
        const int rsub = lane >> 2, c4 = lane & 3;             // this lane's row within a group of 8 / float4 within the 16 columns
        const bool vec_ok = !(p.ldc & 3) && !(reinterpret_cast<uintptr_t>(p.C) & 15) && !(p.c_zs & 3) && !(p.c_tap_stride & 3) &&
                            (!p.addend || (!(p.ld_add & 3) && !(reinterpret_cast<uintptr_t>(p.addend) & 15)));
        int acc = 0, acc_par = 0;
        for (int it_ = 0;; ++it_) {
            const int u = item_unit<HC>(p, it_, pair, npairs, MP, nblocks, total);
            if (u < 0) break;
            const Unit t = decode_unit(p, u, MP, nblocks, crank);
            if (t.KB <= 0) continue;
            mbar_wait(BAR(BAR_T_FULL + acc), acc_par);
            tc_fence_after();
            if (p.dbg && it_ == 0 && crank == 0 && q == 0 && cg == 0 && lane == 0) {
                unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                p.dbg[pair * DBG_STRIDE + 11] = (long long)(gt - gt_start);     // ns until the first accumulator stage was complete
            }
            float* Cb = p.C + t.c_z + (long long)t.ytap * p.c_tap_stride;
            const float* addb = p.addend ? p.addend + t.c_z + (long long)t.ytap * p.c_tap_stride : nullptr;
            const int grow0 = t.m0 + q * 32;
            const int nrows = min(32, t.rows - q * 32);        // <= 0 for padding rows
            const bool atom = p.atomic || (p.split_red && t.sliced);
            const long long crow0 = (long long)grow0 * p.c_mul + p.c_off;
#pragma unroll 1
            for (int cb = cg; cb < GEMM_BN / 16; cb += ncg) {
                const int col0 = cb * 16;
                // addend rows of this block (attention: dQ accumulates onto the decoder-input gradient): loaded BEFORE the
                // accumulator is read, so that their latency runs under the tcgen05.ld and the staging (they used to be four
                // dependent global loads per block on the epilogue's critical path: 25 us of a 42 us launch)
                float4 adv[4];
                const bool add_vec = addb && vec_ok && !atom && t.n0 + col0 + c4 * 4 + 4 <= p.N;
                if (add_vec) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = rsub + 8 * i;
                        adv[i] = rr < nrows ? __ldg(reinterpret_cast<const float4*>(addb + (crow0 + (long long)rr * p.c_mul) * p.ld_add + t.n0 + col0 + c4 * 4))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * GEMM_BN + col0, r);
                tmem_ld_wait();
                if (cb == cb_last) {                           // accumulator stage fully read by this warp: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (crank == 0) mbar_arrive(BAR(BAR_T_EMPTY + acc)); else mbar_arrive_remote(BAR(BAR_T_EMPTY + acc), 0); }
                }
                if (nrows <= 0 || t.n0 + col0 >= p.N) continue;   // warp-uniform
                // lane = row: float4 j of the row goes to slot j ^ ((row >> 1) & 3) (conflict-free both ways)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stage + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                const int gcol = t.n0 + col0 + c4 * 4;
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (p.bias && t.kb0 == 0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (gcol + e < p.N) bv[e] = __ldg(p.bias + gcol + e);
                }
                const bool full = vec_ok && gcol + 4 <= p.N;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = rsub + 8 * i;
                    if (rr >= nrows || gcol >= p.N) continue;
                    const float4 v = *reinterpret_cast<const float4*>(stage + rr * 16 + ((c4 ^ ((rr >> 1) & 3)) << 2));
                    float o[4] = {v.x * p.alpha + bv[0], v.y * p.alpha + bv[1], v.z * p.alpha + bv[2], v.w * p.alpha + bv[3]};
                    float* dst = Cb + (crow0 + (long long)rr * p.c_mul) * p.ldc + gcol;
                    if (full) {
                        if (atom) red_add_v4(dst, o[0], o[1], o[2], o[3]);
                        else {
                            if (addb) {                        // (full && !atom here, i.e. add_vec: loaded above)
                                const float4 a = adv[i];
                                o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
                            }
                            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                            if (p.Chi) {
                                const long long pi = (long long)t.z * p.cp_zs + (crow0 + (long long)rr * p.c_mul) * p.ldcp + gcol;
                                uint2 hh, ll;
                                split4(make_float4(o[0], o[1], o[2], o[3]), hh, ll);
                                *reinterpret_cast<uint2*>(p.Chi + pi) = hh;
                                *reinterpret_cast<uint2*>(p.Clo + pi) = ll;
                            }
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (gcol + e >= p.N) break;
                            if (atom) atomicAdd(dst + e, o[e]);
                            else dst[e] = o[e] + (addb ? __ldg(addb + (crow0 + (long long)rr * p.c_mul) * p.ld_add + gcol + e) : 0.f);
                        }
                    }
                }
                __syncwarp();
            }
            if (++acc == N_ACC) { acc = 0; acc_par ^= 1; }
            if (item_closes_tile<HC>(p, it_, pair, npairs, nblocks)) {
                // every column block of this CTA's 128 rows is in z now (written by the 16 epilogue warps of this CTA):
                // make it visible CTA-wide, then finish the rows -- the MMA warp is already busy with the next row tile
                asm volatile("bar.sync 13, 512;" ::: "memory");
                const float* z0 = p.C + (long long)t.m0 * p.ldc;
                const int wq = (cg << 2) | q;                  // 0..15, any bijection of the 16 warps
                if (p.hc_C == 256) hc_tail_rows<1>(p, sStage, wq, lane, t.m0, t.rows, z0);
                else if (p.hc_C == 512) hc_tail_rows<2>(p, sStage, wq, lane, t.m0, t.rows, z0);
                else hc_tail_rows<4>(p, sStage, wq, lane, t.m0, t.rows, z0);
            }
        }
        if (p.dbg && crank == 0 && q == 0 && cg == 0 && lane == 0) {
            unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            p.dbg[pair * DBG_STRIDE + 12] = (long long)(gt - gt_start);         // ns until this warp had drained its last unit
        }
    };


    // ================================================================ epilogue worker with the conv tail fused in
    // N <= 256: the 16 epilogue warps of a CTA hold complete rows between them (4 warps x 64 columns per 32-row lane
    // quadrant, lane = row).  Two passes over tensor memory give the two-pass row moments of z = acc + bias (partial sums
    // exchanged through the warps' staging buffers, one named barrier per quadrant), a third pass normalises, applies
    // ReLU / dropout and writes z, y, the operand planes and the row statistics: no separate LayerNorm launch, no
    // re-read of z.
    auto epilogue_ln = [&](const int q, const int cg, float* stage) {
        constexpr int ncg = EPI_WARPS / 4;
        const int rsub = lane >> 2, c4 = lane & 3;
        float* qstage = sStage + (q * 512);                     // staging buffer of warp (q, cg = 0); warp w's is at w * 512
        const float inv_n = 1.f / (float)p.N;
        const float inv_keep = p.ln_drop > 0.f ? 1.f / (1.f - p.ln_drop) : 1.f;
        const unsigned long long sd = eff_seed(p.ln_seed, p.ln_step);
        auto quad_sum = [&](float part) {                       // sum over the 4 warps that share this lane quadrant
            stage[lane] = part;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
            float tot = 0.f;
#pragma unroll
            for (int g = 0; g < ncg; ++g) tot += qstage[g * 4 * 512 + lane];
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
            return tot;
        };
        int acc = 0, acc_par = 0;
        for (int it_ = 0;; ++it_) {
            const int u = item_unit<HC>(p, it_, pair, npairs, MP, nblocks, total);
            if (u < 0) break;
            const Unit t = decode_unit(p, u, MP, nblocks, crank);
            if (t.KB <= 0) continue;
            mbar_wait(BAR(BAR_T_FULL + acc), acc_par);
            tc_fence_after();
            const int grow0 = t.m0 + q * 32;
            const int nrows = min(32, t.rows - q * 32);        // <= 0 for padding rows
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * GEMM_BN;
            // ---- pass 1 / 2: mean and centred second moment of this lane's row
            float mean = 0.f, rstd = 1.f;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                float part = 0.f;
#pragma unroll 1
                for (int cb = cg; cb < GEMM_BN / 16; cb += ncg) {
                    const int col0 = cb * 16;
                    if (col0 >= p.N) break;
                    uint32_t r[16];
                    tmem_ld16(tbase + col0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (col0 + j < p.N) {
                            const float v = __uint_as_float(r[j]) * p.alpha + (p.bias ? __ldg(p.bias + col0 + j) : 0.f);
                            part += pass == 0 ? v : (v - mean) * (v - mean);
                        }
                    }
                }
                const float tot = quad_sum(part);
                if (pass == 0) mean = tot * inv_n; else rstd = rsqrtf(tot * inv_n + 1e-12f);
            }
            if (p.ln_stats && cg == 0 && lane < nrows) *reinterpret_cast<float2*>(p.ln_stats + (long long)(grow0 + lane) * 2) = make_float2(mean, rstd);
            // ---- pass 3: transposed, 16-byte wide outputs
#pragma unroll 1
            for (int cb = cg; cb < GEMM_BN / 16; cb += ncg) {
                const int col0 = cb * 16;
                const bool live = nrows > 0 && col0 < p.N;
                uint32_t r[16];
                if (live) { tmem_ld16(tbase + col0, r); tmem_ld_wait(); }
                if (cb + ncg >= GEMM_BN / 16) {                // last block of this warp: hand the accumulator stage back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (crank == 0) mbar_arrive(BAR(BAR_T_EMPTY + acc)); else mbar_arrive_remote(BAR(BAR_T_EMPTY + acc), 0); }
                }
                if (!live) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stage + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                const int gcol = col0 + c4 * 4;
                float bv[4] = {0.f, 0.f, 0.f, 0.f}, gm[4] = {0.f, 0.f, 0.f, 0.f}, bt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) if (gcol + e < p.N) {
                    if (p.bias) bv[e] = __ldg(p.bias + gcol + e);
                    gm[e] = __ldg(p.ln_gamma + gcol + e); bt[e] = __ldg(p.ln_beta + gcol + e);
                }
                const bool full = gcol + 4 <= p.N;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = rsub + 8 * i;
                    const float m_r = __shfl_sync(0xffffffffu, mean, rr), s_r = __shfl_sync(0xffffffffu, rstd, rr);
                    if (rr >= nrows || gcol >= p.N) continue;
                    const float4 v = *reinterpret_cast<const float4*>(stage + rr * 16 + ((c4 ^ ((rr >> 1) & 3)) << 2));
                    const long long row = grow0 + rr;
                    float zv[4] = {v.x * p.alpha + bv[0], v.y * p.alpha + bv[1], v.z * p.alpha + bv[2], v.w * p.alpha + bv[3]};
                    float yv[4], sg[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float uu = (zv[e] - m_r) * s_r * gm[e] + bt[e];
                        sg[e] = sigmoidf_(uu);
                        float a = p.ln_act == 1 ? fmaxf(uu, 0.f) : uu;
                        if (p.ln_drop > 0.f) a *= drop_scale(sd, (unsigned long long)row * p.N + gcol + e, p.ln_drop, inv_keep);
                        yv[e] = a;
                    }
                    if (full) {
                        if (p.C) *reinterpret_cast<float4*>(p.C + row * p.ldc + gcol) = make_float4(zv[0], zv[1], zv[2], zv[3]);
                        *reinterpret_cast<float4*>(p.ln_y + row * p.ln_ldy + gcol) = make_float4(yv[0], yv[1], yv[2], yv[3]);
                        if (p.ln_ysig) *reinterpret_cast<float4*>(p.ln_ysig + row * p.ln_ldys + gcol) = make_float4(sg[0], sg[1], sg[2], sg[3]);
                        if (p.ln_yhi) {
                            uint2 hh, ll;
                            split4(make_float4(yv[0], yv[1], yv[2], yv[3]), hh, ll);
                            *reinterpret_cast<uint2*>(p.ln_yhi + row * p.ln_ldp + gcol) = hh;
                            *reinterpret_cast<uint2*>(p.ln_ylo + row * p.ln_ldp + gcol) = ll;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (gcol + e >= p.N) break;
                            if (p.C) p.C[row * p.ldc + gcol + e] = zv[e];
                            p.ln_y[row * p.ln_ldy + gcol + e] = yv[e];
                            if (p.ln_ysig) p.ln_ysig[row * p.ln_ldys + gcol + e] = sg[e];
                            if (p.ln_yhi) st_split1(p.ln_yhi, p.ln_ylo, row * p.ln_ldp + gcol + e, yv[e]);
                        }
                    }
                }
                __syncwarp();
            }
            if (++acc == N_ACC) { acc = 0; acc_par ^= 1; }
        }
    };

    // Register budget per role (768 threads launch at 80 regs): producer warpgroups grow to 88, the epilogue
    // warpgroup shrinks to 72 and the MMA / copy warpgroup to 40 (512*88 + 128*72 + 128*40 <= 768*80: setmaxnreg can
    // only hand out what the CTA got at launch).
    // (the MMA / copy / fix-up roles come first in the code: they sit on the start-up path of every launch, and straight-line
    //  code after the prologue is what the instruction fetch already has in flight)
    if (warp >= NPW + 4) {
      if (p.dbg && crank == 0 && warp == NPW + 5 && lane == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); p.dbg[pair * DBG_STRIDE + 18] = (long long)(gt_ - gt_start); }
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == NPW + 5) {
        // ================================================================ copy-engine loader (one thread per CTA)
        // Both CTAs of the pair copy their halves with cta_group::2 tensor copies that signal the LEADER's FULL
        // barrier directly; the leader's loader announces the bytes of both (arrive.expect_tx).
        if (lane == 0 && (packed || p.a_tma || p.r_tma)) {
            if (p.dbg && crank == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); p.dbg[pair * DBG_STRIDE + 19] = (long long)(gt_ - gt_start); }
            int slot = 0, par = 1, as = 0, a_par = 1;
            const uint32_t fullA0 = mapa_u32(BAR(BAR_FULL_A), 0), fullB0 = mapa_u32(BAR(BAR_FULL_B), 0);   // leader's barriers
            const int KBI = (p.A.L + GEMM_BK - 1) / GEMM_BK;       // 64-step blocks per batch item (r_tma, split-K)
            for (int it_ = 0;; ++it_) {
                const int u = item_unit<HC>(p, it_, pair, npairs, MP, nblocks, total);
                if (u < 0) break;
                const Unit t = decode_unit(p, u, MP, nblocks, crank);
                if (p.dbg && crank == 0 && it_ == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); p.dbg[pair * DBG_STRIDE + 20] = (long long)(gt_ - gt_start); }
                // packed image rows (128 bytes each) of this CTA's half of stage (nb, kb): ((nb*KB + kb)*2 + crank) * 256
                const int brow0 = (t.nb * t.KB * 2 + (int)crank) * 256;
                // CTA pairs walk the k-blocks of a unit from different starting points (rot): at any moment they ask the
                // L2 for different weight stages instead of all hammering the same 64 KiB
                const int nkb = t.kb1 - t.kb0, rot = !rotate ? 0 : (nkb == t.KB ? pair - fdiv(pair, p.fd_KB) * nkb : pair % nkb);
                for (int j = 0; j < nkb; ++j) {
                    int kb = t.kb0 + j + rot; if (kb >= t.kb1) kb -= nkb;
                    // steps [tb, tb + 64) of batch item `item`: the k-block of an MN-major (row-reduction) operand
                    int item = t.z, tb = kb * GEMM_BK;
                    if (p.dbg && crank == 0 && it_ == 0 && j == 0) {
                        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                        p.dbg[pair * DBG_STRIDE + 14] = (long long)(gt - gt_start);     // ns until the loader issues its first copy
                    }
                    if (p.z_mode != Z_BATCH) { const int g = t.k_begin + kb; item = fdiv(g, p.fd_KBI); tb = (g - item * KBI) * GEMM_BK; }
                    // ------------------------------------------------ A
                    if (p.r_tma) {         // x^T tile = 64 steps x 128 channels [m0, m0+128) as two 64-channel boxes per plane
                        mbar_wait(BAR(BAR_EMPTY_A + as), a_par);
                        const uint32_t bar = fullA0 + 8u * as;
                        const uint32_t dst = smem_u32(sA + as * A_SLOT);
                        if (crank == 0) mbar_arrive_expect_tx(BAR(BAR_FULL_A + as), 2 * A_SLOT);
                        const int ta = tb + p.A.off[t.ytap];
                        tma_load_3d_cg2(dst, &p.tmA_hi, t.m0, ta, item, bar);
                        tma_load_3d_cg2(dst + 8192, &p.tmA_hi, t.m0 + 64, ta, item, bar);
                        tma_load_3d_cg2(dst + A_PLANE, &p.tmA_lo, t.m0, ta, item, bar);
                        tma_load_3d_cg2(dst + A_PLANE + 8192, &p.tmA_lo, t.m0 + 64, ta, item, bar);
                        if (++as == NA_SLOTS) { as = 0; a_par ^= 1; }
                    } else if (p.a_tma) {  // K-major tile: two boxes (hi / lo plane) of 64 channels x 128 rows
                        const int tap = fdiv(kb, p.fd_KBc), cb = kb - tap * t.KBc;
                        mbar_wait(BAR(BAR_EMPTY_A + as), a_par);
                        const uint32_t dst = smem_u32(sA + as * A_SLOT);
                        const int off = p.A.off[tap];
                        if (p.dbg_flags & 128) {               // diagnostics: no A traffic at all
                            if (crank == 0) { mbar_arrive(BAR(BAR_FULL_A + as)); mbar_arrive(BAR(BAR_FULL_A + as)); mbar_arrive(BAR(BAR_FULL_A + as)); }
                        } else {
                            const bool flat = p.a_tma == 1;
                            const bool fix_me = flat && !nofix && needs_fix(t.m0, off, p.A.L, p.fd_L);
                            if (crank == 0) {                  // bytes that will be signalled straight on FULL_A
                                const bool fix_peer = flat && !nofix && needs_fix(t.m0 + GEMM_BM, off, p.A.L, p.fd_L);
                                mbar_arrive_expect_tx(BAR(BAR_FULL_A + as), (fix_me ? 0 : A_SLOT) + (fix_peer ? 0 : A_SLOT));
                            }
                            const int c0 = t.k_begin + cb * GEMM_BK;
                            if (fix_me) {                      // lands locally; the fix-up warp sends this CTA's token
                                const uint32_t bar = BAR(BAR_LAND_A + as);
                                mbar_arrive_expect_tx(bar, A_SLOT);
                                tma_load_2d(dst, &p.tmA_hi, c0, t.m0 + off, bar);
                                tma_load_2d(dst + A_PLANE, &p.tmA_lo, c0, t.m0 + off, bar);
                            } else {
                                const uint32_t bar = fullA0 + 8u * as;
                                if (flat) {
                                    tma_load_2d_cg2(dst, &p.tmA_hi, c0, t.m0 + off, bar);
                                    tma_load_2d_cg2(dst + A_PLANE, &p.tmA_lo, c0, t.m0 + off, bar);
                                } else {
                                    tma_load_3d_cg2(dst, &p.tmA_hi, c0, t.m0, t.z, bar);
                                    tma_load_3d_cg2(dst + A_PLANE, &p.tmA_lo, c0, t.m0, t.z, bar);
                                }
                                if (crank == 0) mbar_arrive(BAR(BAR_FULL_A + as)); else mbar_arrive_remote(BAR(BAR_FULL_A + as), 0);
                            }
                        }
                        if (++as == NA_SLOTS) { as = 0; a_par ^= 1; }
                    }
                    // ------------------------------------------------ B
                    if (p.r_tma || p.b_tma == 3) {             // MN-major: 64 steps x this CTA's 128 of the 256 output columns
                        mbar_wait(BAR(BAR_EMPTY_B + slot), par);
                        const uint32_t bar = fullB0 + 8u * slot;
                        const uint32_t dst = smem_u32(sB + slot * B_SLOT);
                        if (crank == 0) mbar_arrive_expect_tx(BAR(BAR_FULL_B + slot), 2 * B_SLOT);
                        const int tbb = tb + (p.r_tma ? p.Bm.off[t.ytap] : 0), n = t.n0 + (int)crank * GEMM_BNC;
                        tma_load_3d_cg2(dst, &p.tmB_hi, n, tbb, item, bar);
                        tma_load_3d_cg2(dst + 8192, &p.tmB_hi, n + 64, tbb, item, bar);
                        tma_load_3d_cg2(dst + B_PLANE, &p.tmB_lo, n, tbb, item, bar);
                        tma_load_3d_cg2(dst + B_PLANE + 8192, &p.tmB_lo, n + 64, tbb, item, bar);
                        if (++slot == NB_SLOTS) { slot = 0; par ^= 1; }
                    } else if (p.b_tma == 2) {                 // K-major rows [n, n + 128) of item z, 64 channels
                        mbar_wait(BAR(BAR_EMPTY_B + slot), par);
                        const uint32_t bar = fullB0 + 8u * slot;
                        const uint32_t dst = smem_u32(sB + slot * B_SLOT);
                        if (crank == 0) mbar_arrive_expect_tx(BAR(BAR_FULL_B + slot), 2 * B_SLOT);
                        const int n = t.n0 + (int)crank * GEMM_BNC, c0 = t.k_begin + kb * GEMM_BK;
                        tma_load_3d_cg2(dst, &p.tmB_hi, c0, n, t.z, bar);
                        tma_load_3d_cg2(dst + B_PLANE, &p.tmB_lo, c0, n, t.z, bar);
                        if (++slot == NB_SLOTS) { slot = 0; par ^= 1; }
                    } else if (packed) {
                        mbar_wait(BAR(BAR_EMPTY_B + slot), par);
                        if (p.dbg_flags & 2) { if (crank == 0) mbar_arrive(BAR(BAR_FULL_B + slot)); }
                        else {
                            if (crank == 0) mbar_arrive_expect_tx(BAR(BAR_FULL_B + slot), 2 * B_SLOT);
                            tma_load_2d_cg2(smem_u32(sB + slot * B_SLOT), &p.tmB_hi, 0, brow0 + kb * 512, fullB0 + 8u * slot);
                        }
                        if (++slot == NB_SLOTS) { slot = 0; par ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
      } else if (warp == NPW + 6) {
        // ================================================================ item-boundary fix-up (whole warp, both CTAs)
        // Flat conv-style A tiles: rows whose tap-shifted source step falls outside their own batch item were fetched
        // from the neighbouring item.  They are conv padding: such tiles land on a local barrier, the rows are zeroed
        // here (generic-proxy stores + proxy fence) and the bytes are then accounted on the leader's FULL barrier.
        if (p.a_tma == 1) {
            int as = 0; uint32_t land_par = 0;
            for (int it_ = 0;; ++it_) {
                const int u = item_unit<HC>(p, it_, pair, npairs, MP, nblocks, total);
                if (u < 0) break;
                const Unit t = decode_unit(p, u, MP, nblocks, crank);
                int tmod[4];                                   // step within the item of this lane's 4 rows
#pragma unroll
                for (int j = 0; j < 4; ++j) { const int r = t.m0 + lane + 32 * j; tmod[j] = r - fdiv(r, p.fd_L) * p.A.L; }
                const int nkb = t.kb1 - t.kb0, rot = !rotate ? 0 : (nkb == t.KB ? pair - fdiv(pair, p.fd_KB) * nkb : pair % nkb);
                for (int j = 0; j < nkb; ++j) {
                    int kb = t.kb0 + j + rot; if (kb >= t.kb1) kb -= nkb;
                    const int off = p.A.off[fdiv(kb, p.fd_KBc)];
                    if (!nofix && needs_fix(t.m0, off, p.A.L, p.fd_L)) {
                        mbar_wait(BAR(BAR_LAND_A + as), (land_par >> as) & 1u);
                        land_par ^= 1u << as;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int ts = tmod[r] + off;
                            if (ts < 0 || ts >= p.A.L) {
                                uint4* h = reinterpret_cast<uint4*>(sA + as * A_SLOT + (lane + 32 * r) * 128);
                                uint4* l = reinterpret_cast<uint4*>(sA + as * A_SLOT + A_PLANE + (lane + 32 * r) * 128);
#pragma unroll
                                for (int e = 0; e < 8; ++e) { h[e] = make_uint4(0u, 0u, 0u, 0u); l[e] = make_uint4(0u, 0u, 0u, 0u); }
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) { if (crank == 0) mbar_arrive(BAR(BAR_FULL_A + as)); else mbar_arrive_remote(BAR(BAR_FULL_A + as), 0); }
                    }
                    if (++as == NA_SLOTS) as = 0;
                }
            }
        }
        __syncwarp();
      } else if (warp == NPW + 4) {
        // ================================================================ MMA issuer: one thread of the leader CTA
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, GEMM_BN, p.a_mode == A_MNMAJOR, p.b_mode == B_MNMAJOR);
            const uint32_t a_step = (p.a_mode == A_MNMAJOR) ? 2048u : 32u;
            const uint32_t a_lbo = (p.a_mode == A_MNMAJOR) ? 8192u : 16u;
            const uint32_t b_step = (p.b_mode == B_MNMAJOR) ? 2048u : 32u;
            const uint32_t b_lbo = (p.b_mode == B_MNMAJOR) ? 8192u : 16u;
            const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
            int as = 0, a_par = 0, bs = 0, b_par = 0, acc = 0, acc_par = 1;
            long long w_t = 0, w_a = 0, w_b = 0, t_begin = clock64(), nkb = 0;
            for (int it_ = 0;; ++it_) {
                const int u = item_unit<HC>(p, it_, pair, npairs, MP, nblocks, total);
                if (u < 0) break;
                const Unit t = decode_unit(p, u, MP, nblocks, crank);
                if (t.KB <= 0) continue;
                long long c0 = clock64();
                mbar_wait(BAR(BAR_T_EMPTY + acc), acc_par);    // epilogue of the unit that last used this stage is done
                w_t += clock64() - c0;
                tc_fence_after();
                const uint32_t d = tmem_base + acc * GEMM_BN;
                const int nk = t.kb1 - t.kb0;
                for (int j = 0; j < nk; ++j) {
                    c0 = clock64();
                    mbar_wait(BAR(BAR_FULL_A + as), a_par);
                    const long long c1 = clock64();
                    if (p.dbg && nkb == 0) {
                        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                        p.dbg[pair * DBG_STRIDE + 15] = (long long)(gt - gt_start);     // ns until the first A tile had landed
                    }
                    mbar_wait(BAR(BAR_FULL_B + bs), b_par);
                    w_a += c1 - c0; w_b += clock64() - c1; ++nkb;
                    tc_fence_after();
                    if (p.dbg && nkb == 1) {
                        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                        p.dbg[pair * DBG_STRIDE + 10] = (long long)(gt - gt_start);     // ns until the first operands had landed
                    }
                    const uint32_t a_hi = sA_addr + as * A_SLOT, a_lo = a_hi + A_PLANE;
                    const uint32_t b_hi = sB_addr + bs * B_SLOT, b_lo = b_hi + B_PLANE;
#pragma unroll
                    for (int ks = 0; ks < GEMM_BK / 16; ++ks) {
                        const uint64_t dah = make_sdesc(a_hi + ks * a_step, a_lbo, 1024);
                        const uint64_t dal = make_sdesc(a_lo + ks * a_step, a_lbo, 1024);
                        const uint64_t dbh = make_sdesc(b_hi + ks * b_step, b_lbo, 1024);
                        const uint64_t dbl = make_sdesc(b_lo + ks * b_step, b_lbo, 1024);
                        umma2_bf16(d, dah, dbh, idesc, (j | ks) != 0);
                        umma2_bf16(d, dah, dbl, idesc, 1);
                        umma2_bf16(d, dal, dbh, idesc, 1);
                    }
                    umma2_commit_mcast(BAR(BAR_EMPTY_B + bs), 3);   // frees the B slot in both CTAs
                    umma2_commit_mcast(BAR(BAR_EMPTY_A + as), 3);   // frees the A slot in both CTAs
                    if (++bs == NB_SLOTS) { bs = 0; b_par ^= 1; }
                    if (++as == NA_SLOTS) { as = 0; a_par ^= 1; }
                }
                umma2_commit_mcast(BAR(BAR_T_FULL + acc), 3);       // accumulators of both CTAs complete
                if (++acc == N_ACC) { acc = 0; acc_par ^= 1; }
            }
            if (p.dbg) {
                long long* o = p.dbg + pair * DBG_STRIDE;
                unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                o[0] = clock64() - t_begin; o[1] = w_t; o[2] = w_a; o[3] = w_b; o[4] = nkb;
                o[5] = (long long)(gt - gt_start);      // ns from kernel entry to the end of the MMA issue loop
            }
        }
        __syncwarp();
      }
    } else if (warp < NPW) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
        if (copy_fed) {
            if (p.ln_fuse) epilogue_ln(warp & 3, warp >> 2, sStage + warp * 512);
            else epilogue(warp & 3, warp >> 2, EPI_WARPS / 4, sStage + warp * 512);
        }
        // ================================================================ producers (both CTAs)
        // Each thread owns NCH chunks (8 consecutive elements) of every operand tile.  Global loads run one k-block
        // ahead of the shared-memory stores (register double buffer); address arithmetic is hoisted out of the chunk
        // loop: row offsets are set up once per tap (conv-style A) or advanced incrementally (row-reduction operands).
        constexpr int NCH = 1024 / (NPW * 32);
        uint32_t off_a[NCH], off_b[NCH];                       // smem offsets: K-major [128][64 k] / MN-major [64 k][128]
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int rl = (tid >> 3) + (NPW * 4) * i, chunk = tid & 7;
            const uint32_t ok_ = rl * 128 + ((chunk ^ (rl & 7)) << 4);
            const int kl = (tid >> 4) + (NPW * 2) * i, j = tid & 15;
            const uint32_t omn = (j >> 3) * 8192 + (kl >> 3) * 1024 + (kl & 7) * 128 + (((j & 7) ^ (kl & 7)) << 4);
            off_a[i] = a_k ? ok_ : omn;
            off_b[i] = (p.b_mode == B_KMAJOR) ? ok_ : omn;
        }
        const bool a_split = p.A.hi != nullptr, b_split = p.Bm.hi != nullptr;
        int a_slot = 0, a_par = 1, b_slot = 0, b_par = 1;      // ring cursors: parity to wait for on the EMPTY barriers
        for (int u = copy_fed ? total : pair; u < total; u += npairs) {
            const Unit t = decode_unit(p, u, MP, nblocks, crank);
            if (t.KB <= 0) continue;
            const int ntl = a_k ? p.ntaps : 1;

            // ---- A load stream (runs one k-block ahead of the store stream)
            int la_tap = 0, la_cb = 0;
            long long a_cur[NCH]; bool a_val[NCH];             // element offsets of this thread's rows for the current tap
            int a_t[NCH]; long long a_base[NCH]; bool a_ok[NCH];
            int ar_b[NCH], ar_t[NCH], ar_r[NCH], br_b[NCH], br_t[NCH], br_r[NCH];
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (a_k) {
                    const int g = t.m0 + (tid >> 3) + (NPW * 4) * i;
                    a_ok[i] = g < p.M;
                    const int b = g / p.A.L;
                    a_t[i] = g - b * p.A.L;
                    a_base[i] = (long long)b * p.A.Ls;
                } else {
                    ar_r[i] = t.k_begin + (tid >> 4) + (NPW * 2) * i;
                    ar_b[i] = ar_r[i] / p.A.L;
                    ar_t[i] = ar_r[i] - ar_b[i] * p.A.L;
                }
                if (p.b_mode == B_MNMAJOR) {
                    br_r[i] = t.k_begin + (tid >> 4) + (NPW * 2) * i;
                    br_b[i] = br_r[i] / p.Bm.L;
                    br_t[i] = br_r[i] - br_b[i] * p.Bm.L;
                }
            }
            auto a_set_tap = [&](int tap) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    const int ts = a_t[i] * p.A.mul + p.A.off[tap];
                    a_val[i] = a_ok[i] && ts >= 0 && ts < p.A.Ls;
                    a_cur[i] = (a_base[i] + ts) * p.A.ld + (tid & 7) * 8;
                }
            };
            if (a_k) a_set_tap(0);
            auto load_a = [&](float (&v)[NCH][8]) {              // next k-block of the stream: global -> registers
                if ((p.dbg_flags & 1) || p.a_tma) return;
                if (a_k) {
                    const int kk = t.k_begin + la_cb * GEMM_BK;
                    const int c = kk + (tid & 7) * 8;
#pragma unroll
                    for (int i = 0; i < NCH; ++i) load_chunk(p.A, t.a_z, a_cur[i] + kk, a_val[i], c, t.k_end, v[i]);
                    if (++la_cb == t.KBc) { la_cb = 0; if (++la_tap < ntl) a_set_tap(la_tap); }
                } else {
                    const int m = t.m0 + (tid & 15) * 8;
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        const int ts = ar_t[i] * p.A.mul + p.A.off[t.ytap];
                        const bool ok = ts >= 0 && ts < p.A.Ls && ar_r[i] < t.k_end;
                        load_chunk(p.A, t.a_z, ((long long)ar_b[i] * p.A.Ls + ts) * p.A.ld + m, ok, m, p.M, v[i]);
                        ar_r[i] += GEMM_BK; ar_t[i] += GEMM_BK;
                        while (ar_t[i] >= p.A.L) { ar_t[i] -= p.A.L; ++ar_b[i]; }
                    }
                }
            };
            const int nbase = t.n0 + (int)crank * GEMM_BNC;    // this CTA stages B rows/columns [nbase, nbase + 128)
            auto load_b = [&](int kb, float (&v)[NCH][8]) {      // B from activations (attention / weight-gradient products)
                if (p.b_mode == B_KMAJOR) {
                    const int c = t.k_begin + (kb - (kb / t.KBc) * t.KBc) * GEMM_BK + (tid & 7) * 8;
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        const int n = nbase + (tid >> 3) + (NPW * 4) * i;
                        load_chunk(p.Bm, t.b_z, (long long)n * p.Bm.ld + c, n < p.N, c, t.k_end, v[i]);
                    }
                } else {
                    const int tap = a_k ? kb / t.KBc : t.ytap;
                    const int n = nbase + (tid & 15) * 8;
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        const int ts = br_t[i] * p.Bm.mul + p.Bm.off[tap];
                        const bool ok = ts >= 0 && ts < p.Bm.Ls && br_r[i] < t.k_end;
                        load_chunk(p.Bm, t.b_z, ((long long)br_b[i] * p.Bm.Ls + ts) * p.Bm.ld + n, ok, n, p.N, v[i]);
                        br_r[i] += GEMM_BK; br_t[i] += GEMM_BK;
                        while (br_t[i] >= p.Bm.L) { br_t[i] -= p.Bm.L; ++br_b[i]; }
                    }
                }
            };
            // one k-block: (convert +) store the A tile held in registers, then B for activation x activation products
            float wb[NCH][8];
            auto emit = [&](int kb, float (&v)[NCH][8]) {
                if (!p.a_tma) {
                    mbar_wait(BAR(BAR_EMPTY_A + a_slot), a_par);
                    uint8_t* hi = sA + a_slot * A_SLOT;
                    if (!(p.dbg_flags & 1)) {
#pragma unroll
                        for (int i = 0; i < NCH; ++i) store_chunk(a_split, hi, hi + A_PLANE, off_a[i], v[i]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { if (crank == 0) mbar_arrive(BAR(BAR_FULL_A + a_slot)); else mbar_arrive_remote(BAR(BAR_FULL_A + a_slot), 0); }
                    if (++a_slot == NA_SLOTS) { a_slot = 0; a_par ^= 1; }
                }
                if (!packed) {                                 // wb was loaded one k-block ahead
                    mbar_wait(BAR(BAR_EMPTY_B + b_slot), b_par);
                    uint8_t* hi = sB + b_slot * B_SLOT;
#pragma unroll
                    for (int i = 0; i < NCH; ++i) store_chunk(b_split, hi, hi + B_PLANE, off_b[i], wb[i]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { if (crank == 0) mbar_arrive(BAR(BAR_FULL_B + b_slot)); else mbar_arrive_remote(BAR(BAR_FULL_B + b_slot), 0); }
                    if (++b_slot == NB_SLOTS) { b_slot = 0; b_par ^= 1; }
                    if (kb + 1 < t.KB) load_b(kb + 1, wb);
                }
            };
            float v0[NCH][8], v1[NCH][8];
            load_a(v0);
            if (!packed) load_b(0, wb);
            for (int kb = 0; kb < t.KB; kb += 2) {
                if (kb + 1 < t.KB) load_a(v1);
                emit(kb, v0);
                if (kb + 1 < t.KB) {
                    if (kb + 2 < t.KB) load_a(v0);
                    emit(kb + 1, v1);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        if (!copy_fed) epilogue(warp & 3, 0, 1, sStage + (warp & 3) * 512);
    }

    tc_fence_before();
    if (p.dbg && crank == 0 && tid == 256) {           // an epilogue-free producer thread: time until its role loop ended
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[pair * DBG_STRIDE + 6] = (long long)(gt - gt_start);
    }
    cluster_sync_all();                                // the partner's smem/barriers stay alive until both are done
    if (p.dbg && crank == 0 && tid == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[pair * DBG_STRIDE + 7] = (long long)(gt - gt_start);     // ns from kernel entry to after the final cluster barrier
        p.dbg[pair * DBG_STRIDE + 8] = (long long)gt_start;            // absolute entry / exit times: launch skew between the pairs
        p.dbg[pair * DBG_STRIDE + 13] = (long long)gt;
    }
    if (warp == NPW + 4) tmem_dealloc2<N_ACC * GEMM_BN>(tmem_base);
}

}  // namespace oph
