// C ABI of libophelia_sm100.so: composes the tcgen05 implicit-GEMM core with the row-wise kernels into the
// reference's operators (modules.conv1d / hc / conv1d_transpose / embed, networks.Attention, losses, Adam).
#include "../../include/ophelia_b200.h"
#include "rowwise.cuh"
#include "gemm_tc.cuh"
#include "arstep.cuh"
#include "attn_fused.cuh"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace oph;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
int g_gemm_dbg_flags = 0;
int g_gemm_dbg_flags_host = 0;     // all bits as passed to oph_gemm_debug_flags (host-side switches)
bool g_hcb_two = true;
int g_hcb_depth = 0;               // ring depth of the highway-backward row kernel (tunable through oph_gemm_debug_flags bits 8..10)
long long* g_gemm_dbg = nullptr;   // optional device buffer [74][24] for in-kernel wait-cycle counters
int g_dbg_slots = 0, g_dbg_next = 0;   // > 0: g_gemm_dbg is a ring of [slots][74][24]; every launch takes the next slot
std::vector<std::string> g_dbg_desc;   // shape of the launch that owns each slot

// Optional per-launch timing of the GEMM core (bench.py roofline): CUDA events around every launch, by tag.
struct ProfRec { cudaEvent_t e0, e1; int tag; double flops; char desc[112]; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

// Optional side stream for weight-gradient GEMMs (oph_wgrad_stream): they are off the critical path of the backward
// chain, so they fork after the row-wise backward kernel of their layer and overlap the input-gradient chain.
thread_local cudaStream_t g_wgrad_stream = nullptr;
thread_local bool g_wgrad_fork = false;
thread_local cudaEvent_t g_fork_ev[16];
thread_local int g_fork_n = 0, g_fork_i = 0;

// Per-speaker channel gates for the NEXT oph_hc_fwd / oph_hc_bwd call of this host thread (oph_lcc_context); consumed by it.
thread_local LccGate g_lcc = {nullptr, nullptr, 0, nullptr};
LccGate take_lcc(int L) {
    LccGate r = g_lcc;
    r.L = L;
    g_lcc = LccGate{nullptr, nullptr, 0, nullptr};
    return r;
}

cudaStream_t fork_wgrad(cudaStream_t main) {
    if (!g_wgrad_fork || g_wgrad_stream == main) return main;
    if (g_fork_n < 16) { cudaEventCreateWithFlags(&g_fork_ev[g_fork_n], cudaEventDisableTiming); ++g_fork_n; g_fork_i = g_fork_n - 1; }
    else g_fork_i = (g_fork_i + 1) % 16;
    cudaEvent_t ev = g_fork_ev[g_fork_i];
    if (cudaEventRecord(ev, main) != cudaSuccess || cudaStreamWaitEvent(g_wgrad_stream, ev, 0) != cudaSuccess) {
        cudaGetLastError();
        return main;
    }
    return g_wgrad_stream;
}

// Every kernel of the library is launched with programmatic stream serialization (PDL): a kernel's blocks may become
// resident while its predecessor in the stream drains and run their prologue; `pdl_grid_sync()` at the top of each
// kernel (griddepcontrol.wait) orders their first global-memory access after the predecessor's completion.
bool g_use_pdl = false;   // measured on B200: no gain inside CUDA graphs (7.60 vs 7.57 ms per step), kept as an option
struct launch_cfg {
    dim3 grid, block; size_t smem; cudaStream_t st;
    launch_cfg(dim3 g, dim3 b, size_t s, cudaStream_t stream) : grid(g), block(b), smem(s), st(stream) {}
    template <typename... KA, typename... A>
    void operator()(void (*kernel)(KA...), A&&... args) const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
    }
};

// CUDA events around one launch while profiling is on (oph_profile_begin/end); `work` = algorithmic FLOPs (GEMM tags) or
// algorithmic bytes (row-wise tags)
struct ProfScope {
    ProfRec rec{}; bool on = false; cudaStream_t st;
    ProfScope(int tag, double work, cudaStream_t stream, const char* desc = "") : st(stream) {
        if (!g_prof_on) return;
        snprintf(rec.desc, sizeof(rec.desc), "%s", desc);
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs != cudaStreamCaptureStatusNone) return;
        on = true;
        cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1);
        rec.tag = tag; rec.flops = work;
        cudaEventRecord(rec.e0, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(rec.e1, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(rec);
    }
};

int fail(int code, const char* fmt, const char* detail = "") {   // fmt contains exactly one %s
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
// ---- trap record of the bounded barrier waits (oph_ptx.cuh): 4 words of mapped pinned host memory per process
unsigned long long* g_trap_host = nullptr;
bool g_trap_armed[64] = {};
void ensure_trap_record(cudaStream_t st = nullptr) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || g_trap_armed[dev]) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;       // allocation + symbol copy are illegal inside a capture:
    if (st && cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return;   // arm at a later launch
    g_trap_armed[dev] = true;
    if (!g_trap_host) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&g_trap_host), 4 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            g_trap_host = nullptr; cudaGetLastError(); return;
        }
        memset(g_trap_host, 0, 4 * sizeof(unsigned long long));
    }
    unsigned long long* d = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), g_trap_host, 0) != cudaSuccess ||
        cudaMemcpyToSymbol(g_oph_trap_rec, &d, sizeof(d)) != cudaSuccess) cudaGetLastError();
}
// "block 17 of 148, warp 20 (GEMM role: MMA issuer) waited on barrier 3 (FULL_B[0]) for parity 1"
int describe_trap(char* out, size_t cap) {
    if (!g_trap_host || g_trap_host[0] != OPH_TRAP_MAGIC) return 0;
    const unsigned block = (unsigned)(g_trap_host[1] >> 32), thread = (unsigned)g_trap_host[1];
    const unsigned bar = (unsigned)(g_trap_host[2] >> 32), parity = (unsigned)g_trap_host[2];
    const unsigned grid = (unsigned)(g_trap_host[3] >> 32), bdim = (unsigned)g_trap_host[3];
    const unsigned warp = thread >> 5, idx = (bar & 1023u) >> 3;
    const char* role = "";
    char name[48] = "";
    if (bdim == (unsigned)GEMM_THREADS) {          // gemm_bf16x3_kernel: role by warp, barrier by its index in the block
        role = warp < (unsigned)NPW ? " (GEMM: producer / epilogue warp)" : warp < (unsigned)NPW + 4 ? " (GEMM: epilogue warp)" :
               warp == (unsigned)NPW + 4 ? " (GEMM: MMA issuer)" : warp == (unsigned)NPW + 5 ? " (GEMM: copy-engine loader)" :
               warp == (unsigned)NPW + 6 ? " (GEMM: item-boundary fix-up)" : " (GEMM)";
        static const struct { int first, n; const char* nm; } tab[] = {
            {BAR_FULL_A, NA_SLOTS, "FULL_A"}, {BAR_EMPTY_A, NA_SLOTS, "EMPTY_A"}, {BAR_FULL_B, NB_SLOTS, "FULL_B"},
            {BAR_EMPTY_B, NB_SLOTS, "EMPTY_B"}, {BAR_LAND_A, NA_SLOTS, "LAND_A"}, {BAR_T_FULL, N_ACC, "T_FULL"}, {BAR_T_EMPTY, N_ACC, "T_EMPTY"}};
        for (const auto& t : tab)
            if ((int)idx >= t.first && (int)idx < t.first + t.n) snprintf(name, sizeof(name), " = %s[%d]", t.nm, (int)idx - t.first);
    }
    return snprintf(out, cap, "; a bounded barrier wait trapped: block %u of %u, warp %u%s waited on the mbarrier at shared address 0x%x (index %u%s) for parity %u",
                    block, grid, warp, role, bar, idx, name, parity);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        int n = snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        if (n > 0 && n < (int)sizeof(g_err)) describe_trap(g_err + n, sizeof(g_err) - n);
        return OPH_ECUDA;
    }
    return OPH_OK;
}
#define OPH_TRY(expr) do { int _rc = (expr); if (_rc != OPH_OK) return _rc; } while (0)

inline cudaStream_t S(oph_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- TMA tensor maps (driver entry point resolved at run time: no link dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        cudaGetLastError();
    }
    return fn;
}
// bf16 plane [items][L][ld] viewed as (channels, time, item); box = 64 channels x box_rows steps, SWIZZLE_128B, zero fill
bool make_plane_tmap(CUtensorMap* out, const unsigned short* base, int C, int L, int items, long long ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) || (ld & 7)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)items};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<unsigned short*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// bf16 plane [rows][ld] viewed as (channels, rows); box = 64 channels x 128 rows, SWIZZLE_128B, zero fill
bool make_plane_tmap2d(CUtensorMap* out, const unsigned short* base, int C, long long rows, long long ld) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) || (ld & 7)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)GEMM_BM};
    const cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<unsigned short*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool g_use_tma = true;
bool g_use_split = true;
bool g_use_ln_fuse = false;  // measured: 7.02 vs 7.04 ms per training step, but 0.71 vs 0.61 ms per autoregressive frame step
bool g_use_attn_fuse = true; // networks.Attention forward as one kernel (oph_gemm_debug_flags bit 524288 turns it off)
bool g_use_hc_fuse = true;   // highway tail in the epilogue phase of the conv GEMM (oph_gemm_debug_flags bit 131072 turns it off)

int launch_gemm(GemmArgs& a, int zdim, cudaStream_t st) {
    ensure_trap_record(st);
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(gemm_bf16x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_bf16x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(gemm)");
        attr_done = true;
    }
    if (a.M <= 0 || a.N <= 0 || a.Kc <= 0) return fail(OPH_EINVAL, "gemm: empty problem%s");
    if ((a.A.ld & (a.A.hi ? 7 : 3)) || (a.b_mode != B_PACKED && (a.Bm.ld & (a.Bm.hi ? 7 : 3))))
        return fail(OPH_EINVAL, "gemm: row strides must be multiples of 4 (fp32) / 8 (bf16 planes)%s");
    if (a.ytaps < 1) a.ytaps = 1;
    a.zdim = zdim < 1 ? 1 : zdim;
    // conv-style A already split into planes: feed it with TMA tensor copies (flat tiles, padding = out-of-range / fix-up)
    a.a_tma = 0; a.b_tma = 0; if (!a.r_tma) a.items = 1;
    if (g_use_tma && a.a_mode == A_KMAJOR && a.A.hi && a.A.mul == 1 && a.A.L == a.A.Ls && a.z_mode == Z_NONE &&
        a.M % a.A.L == 0) {
        if (make_plane_tmap2d(&a.tmA_hi, a.A.hi, a.Kc, a.M, a.A.ld) && make_plane_tmap2d(&a.tmA_lo, a.A.lo, a.Kc, a.M, a.A.ld)) {
            a.a_tma = 1; a.items = a.M / a.A.L;
        }
    }
    // batched activation x activation products (attention) with both operands as planes: every tile comes from the copy
    // engines; the planes of item z start z * rows * ld elements after those of item 0 (contiguous items)
    if (g_use_tma && a.z_mode == Z_BATCH && a.A.hi && a.Bm.hi && !a.r_tma && a.ytaps == 1 && a.ntaps == 1) {
        if (a.a_mode == A_KMAJOR) {
            bool ok = make_plane_tmap(&a.tmA_hi, a.A.hi, a.Kc, a.M, a.zdim, a.A.ld, GEMM_BM) &&
                      make_plane_tmap(&a.tmA_lo, a.A.lo, a.Kc, a.M, a.zdim, a.A.ld, GEMM_BM);
            if (ok && a.b_mode == B_KMAJOR) {
                ok = make_plane_tmap(&a.tmB_hi, a.Bm.hi, a.Kc, a.N, a.zdim, a.Bm.ld, GEMM_BM) &&
                     make_plane_tmap(&a.tmB_lo, a.Bm.lo, a.Kc, a.N, a.zdim, a.Bm.ld, GEMM_BM);
                if (ok) { a.a_tma = 2; a.b_tma = 2; }
            } else if (ok && a.b_mode == B_MNMAJOR) {
                ok = make_plane_tmap(&a.tmB_hi, a.Bm.hi, a.N, a.Kc, a.zdim, a.Bm.ld, GEMM_BK) &&
                     make_plane_tmap(&a.tmB_lo, a.Bm.lo, a.N, a.Kc, a.zdim, a.Bm.ld, GEMM_BK);
                if (ok) { a.a_tma = 2; a.b_tma = 3; }
            }
        } else if (a.a_mode == A_MNMAJOR && a.b_mode == B_MNMAJOR) {
            const int R = a.Kc;                                            // reduction steps per item
            if (make_plane_tmap(&a.tmA_hi, a.A.hi, a.M, R, a.zdim, a.A.ld, GEMM_BK) && make_plane_tmap(&a.tmA_lo, a.A.lo, a.M, R, a.zdim, a.A.ld, GEMM_BK) &&
                make_plane_tmap(&a.tmB_hi, a.Bm.hi, a.N, R, a.zdim, a.Bm.ld, GEMM_BK) && make_plane_tmap(&a.tmB_lo, a.Bm.lo, a.N, R, a.zdim, a.Bm.ld, GEMM_BK)) {
                a.r_tma = 1; a.items = a.zdim; a.prof_k = R; a.Kc = cdiv(R, GEMM_BK);   // Kc now counts k-blocks
                a.A.off[0] = a.Bm.off[0] = 0;
            }
        }
    }
    // operands that still go through the producer warps as planes are copied in 16-byte chunks
    const bool a_prod = !a.a_tma && !a.r_tma, b_prod = a.b_mode != B_PACKED && !a.b_tma && !a.r_tma;
    if ((a_prod && a.A.hi && a.a_mode == A_KMAJOR && (a.Kc & 7)) || (a_prod && a.A.hi && a.a_mode == A_MNMAJOR && (a.M & 7)) ||
        (b_prod && a.Bm.hi && a.b_mode == B_KMAJOR && (a.Kc & 7)) || (b_prod && a.Bm.hi && a.b_mode == B_MNMAJOR && (a.N & 7)))
        return fail(OPH_EINVAL, "gemm: plane operands need channel counts that are multiples of 8%s");
    if ((a_prod && !a.A.hi && !a.A.ptr) || (b_prod && !a.Bm.hi && !a.Bm.ptr))
        return fail(OPH_EINVAL, "gemm: an operand has neither an fp32 view nor usable planes%s");
    if (a.b_mode == B_PACKED) {   // packed weight image as 128-byte rows: one 256-row box = one CTA's half of a stage
        const size_t rows = (size_t)cdiv(a.N, GEMM_BN) * a.ntaps * cdiv(a.Kc, GEMM_BK) * (B_STAGE / 128);
        EncodeTiledFn fn = encode_tiled_fn();
        const cuuint64_t dims[2] = {64, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, 256};
        const cuuint32_t estr[2] = {1, 1};
        if (!fn || (reinterpret_cast<uintptr_t>(a.Bpacked) & 127) ||
            fn(&a.tmB_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a.Bpacked), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(OPH_ECUDA, "gemm: cannot encode the tensor map of the packed weight image%s");
    }
    a.dbg = g_gemm_dbg;
    int dbg_slot = -1;
    if (g_gemm_dbg && g_dbg_slots > 0) { dbg_slot = g_dbg_next++ % g_dbg_slots; a.dbg = g_gemm_dbg + (size_t)dbg_slot * GEMM_MAX_PAIRS * DBG_STRIDE; }
    a.dbg_flags = g_gemm_dbg_flags;
    // persistent CTA pairs: work units = (pair of 128-row tiles) x (256-column block) x tap x z slice
    // conv tail in the epilogue: whole rows in one work unit (N <= 256), both operands from the copy engines
    a.ln_fuse = (a.ln_gamma && a.ln_y && g_use_ln_fuse && a.a_tma == 1 && a.b_mode == B_PACKED && a.N <= GEMM_BN && !a.atomic &&
                 !a.addend && !a.Chi && a.c_mul == 1 && a.c_off == 0 && a.z_mode == Z_NONE && a.ytaps == 1 &&
                 !(a.ldc & 3) && !(a.ln_ldy & 3) && !(a.ln_ldys & 3) && !(a.ln_ldp & 3)) ? 1 : 0;
    // highway tail in the same launch: only for copy-fed conv products whose epilogue runs on the 16 producer warps
    if (a.hc_fused && !(a.a_tma == 1 && a.b_mode == B_PACKED && a.z_mode == Z_NONE && a.ytaps == 1 && !a.atomic && !a.addend &&
                        !a.Chi && a.c_mul == 1 && a.c_off == 0 && !a.ln_fuse)) a.hc_fused = 0;
    const long long units = (long long)cdiv(cdiv(a.M, GEMM_BM), 2) * cdiv(a.N, GEMM_BN) * a.ytaps * a.zdim;
    // RED outputs fed by the copy engines: split the units of the last (partial) round into k-block slices so that it
    // costs a fraction of a round; s minimises rounds x (slice length + 1 k-block of per-item overhead)
    long long items = units;
    a.split_from = (int)units; a.split_s = 1; a.split_red = 0;
    const bool can_red = a.atomic && !a.bias;
    // plain outputs the caller allows us to pre-zero (a.zero_bytes > 0): the slices of the remainder units RED into zeros
    // (off by default, flag 16384: the arrival order of the REDs would make FORWARD results vary in the last bits from run
    // to run, and with the side streams the idle SMs of a partial round are filled by other kernels anyway)
    // Launches that fill at most half of the CTA pairs (the windowed Attention / AudioDec pass of the incremental
    // autoregressive route: 4-8 units) are always split, but into TWO slices only: 0 + a + b does not depend on the arrival
    // order of the two REDs, so these results stay bit-reproducible.
    const bool small = units * 2 <= GEMM_MAX_PAIRS;
    const bool can_zero = ((g_gemm_dbg_flags_host & 16384) || small) && !a.atomic && a.zero_bytes > 0 && !a.addend && !a.Chi && a.c_mul == 1 &&
                          !a.ln_fuse;      // (the fused conv tail normalises whole rows of the accumulator: no partial sums)
    const int max_slices = (g_gemm_dbg_flags_host & 16384) || can_red ? 1 << 20 : 2;
    if (g_use_split && !a.hc_fused && a.a_tma == 1 && a.b_mode == B_PACKED && (can_red || can_zero) && a.z_mode == Z_NONE && a.ytaps == 1) {
        const int P = GEMM_MAX_PAIRS, KB = a.ntaps * cdiv(a.Kc, GEMM_BK);
        const int rem = (int)(units % P);
        if (rem > 0) {
            int best_s = 1; long long best = (long long)KB + 1;
            for (int sl = 2; sl <= KB / 4 && sl <= max_slices; ++sl) {
                const long long cost = (long long)cdiv(rem * sl, P) * (cdiv(KB, sl) + 1);
                if (cost < best) { best = cost; best_s = sl; }
            }
            const long long zero_cost = can_zero ? 2 + (long long)(a.zero_bytes / 4000000) : 0;   // ~k-block times of the memset
            if (best_s > 1 && best + zero_cost + 4 <= (long long)KB + 1) {
                if (can_zero) {
                    if (cudaMemsetAsync(a.C, 0, a.zero_bytes, st) != cudaSuccess) return check_launch("cudaMemsetAsync(z)");
                    a.split_red = 1;
                }
                a.split_from = (int)(units - rem); a.split_s = best_s; items = a.split_from + (long long)rem * best_s;
            }
        }
    }
    int pairs = (int)(items < GEMM_MAX_PAIRS ? items : GEMM_MAX_PAIRS);
    if (a.hc_fused) {                                   // super-units: one per row-tile pair
        const int mp = cdiv(cdiv(a.M, GEMM_BM), 2);
        pairs = mp < GEMM_MAX_PAIRS ? mp : GEMM_MAX_PAIRS;
    }
    dim3 grid(2 * pairs);
    {   // the launch constants the kernel divides by (same formulas as in the kernel)
        const int KBc = cdiv(a.Kc, GEMM_BK);
        a.fd_MP = make_fastdiv(cdiv(cdiv(a.M, GEMM_BM), 2)); a.fd_nb = make_fastdiv(cdiv(a.N, GEMM_BN)); a.fd_yt = make_fastdiv(a.ytaps);
        a.fd_ss = make_fastdiv(a.split_s); a.fd_L = make_fastdiv(a.A.L); a.fd_KBc = make_fastdiv(KBc);
        a.fd_KB = make_fastdiv((a.a_mode == A_KMAJOR ? a.ntaps : 1) * KBc); a.fd_KBI = make_fastdiv(cdiv(a.A.L, GEMM_BK));
    }
    {
        char desc[112] = "";
        if (g_prof_on || dbg_slot >= 0)
            snprintf(desc, sizeof(desc), "M=%d N=%d K=%d taps=%d ytaps=%d z=%d units=%lld items=%lld pairs=%d split=%d atma=%d btma=%d rtma=%d hc=%d", a.M, a.N,
                     a.prof_k ? a.prof_k : a.Kc, a.ntaps, a.ytaps, a.zdim, units, items, pairs, a.split_s, a.a_tma, a.b_tma, a.r_tma, a.hc_fused);
        if (dbg_slot >= 0) { if ((int)g_dbg_desc.size() <= dbg_slot) g_dbg_desc.resize(dbg_slot + 1); g_dbg_desc[dbg_slot] = std::string("tag=") + std::to_string(a.tag) + " " + desc; }
        ProfScope ps(a.tag, 2.0 * a.M * a.N * (a.prof_k ? a.prof_k : a.Kc) * (a.a_mode == A_KMAJOR ? a.ntaps : a.ytaps) * (a.z_mode == Z_BATCH ? a.zdim : 1), st, desc);
        if (a.hc_fused) launch_cfg(grid, GEMM_THREADS, GEMM_SMEM, st)(gemm_bf16x3_kernel<true>, a);
        else launch_cfg(grid, GEMM_THREADS, GEMM_SMEM, st)(gemm_bf16x3_kernel<false>, a);
    }
    return check_launch("gemm_bf16x3_kernel");
}

GemmArgs blank() {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.ntaps = 1; a.ytaps = 1; a.c_mul = 1; a.alpha = 1.f;
    a.A.L = a.A.Ls = 1; a.A.mul = 1; a.Bm.L = a.Bm.Ls = 1; a.Bm.mul = 1;
    return a;
}

// ------------------------------------------------------------------------------------------------ packing
struct PackArgs {
    const float* w;
    int ntaps, tap_idx[3];
    long long s_tap, s_c, s_n;
    int Cvalid, Nvalid;
    uint8_t* out;
};

__global__ void pack_kernel(const PackArgs p, long long total) {
    pdl_grid_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int KBc = (p.Cvalid + GEMM_BK - 1) / GEMM_BK, KB = p.ntaps * KBc;
    const int nl = (int)(idx & 255);
    const int chunk = (int)((idx >> 8) & 7);
    const long long rest = idx >> 11;
    const int kb = (int)(rest % KB);
    const int nb = (int)(rest / KB);
    const int tap = kb / KBc, cb = kb - tap * KBc;
    const int n = nb * GEMM_BN + nl;
    const int c0 = cb * GEMM_BK + chunk * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c0 + e;
        v[e] = (n < p.Nvalid && c < p.Cvalid) ? p.w[p.tap_idx[tap] * p.s_tap + c * p.s_c + n * p.s_n] : 0.f;
    }
    // stage image = [CTA 0: hi | lo][CTA 1: hi | lo]; each CTA of the pair stages 128 of the 256 rows
    uint8_t* img = p.out + ((size_t)nb * KB + kb) * B_STAGE + (size_t)(nl >> 7) * B_SLOT;
    const int nr = nl & 127;
    store_split(img, img + B_PLANE, nr * 128 + ((chunk ^ (nr & 7)) << 4), v);
}

size_t pack_image_bytes(int ntaps, int Cvalid, int Nvalid) {
    return (size_t)cdiv(Nvalid, GEMM_BN) * ntaps * cdiv(Cvalid, GEMM_BK) * B_STAGE;
}

int pack_image(const float* w, int ntaps, const int* tap_idx, long long s_tap, long long s_c, long long s_n,
               int Cvalid, int Nvalid, void* out, cudaStream_t st) {
    PackArgs p;
    p.w = w; p.ntaps = ntaps;
    for (int i = 0; i < 3; ++i) p.tap_idx[i] = i < ntaps ? tap_idx[i] : 0;
    p.s_tap = s_tap; p.s_c = s_c; p.s_n = s_n; p.Cvalid = Cvalid; p.Nvalid = Nvalid;
    p.out = reinterpret_cast<uint8_t*>(out);
    const long long total = (long long)(pack_image_bytes(ntaps, Cvalid, Nvalid) / B_STAGE) * 2048;
    launch_cfg((unsigned)((total + 255) / 256), 256, 0, st)(pack_kernel, p, total);
    return check_launch("pack_kernel");
}

// ---- batched packing: ONE launch re-packs every conv kernel of a model after the optimiser step
struct PackJob { PackArgs a; long long total; long long first_block; };

__global__ void __launch_bounds__(256) pack_batch_kernel(const PackJob* __restrict__ jobs, int njobs) {
    pdl_grid_sync();
    // one block = one stage (256 image rows x 64 channels = 2048 (row, 8-channel chunk) items, 8 per thread): the job
    // lookup and the stage decode are paid once per block
    __shared__ uint4 s_t[2][2][32][8];                  // [buffer][hi / lo][row][chunk]
    int jl = 0, jh = njobs - 1;                         // last job whose first block is <= blockIdx.x
    while (jl < jh) {
        const int mid = (jl + jh + 1) >> 1;
        if (jobs[mid].first_block <= (long long)blockIdx.x) jl = mid; else jh = mid - 1;
    }
    const PackJob& j = jobs[jl];
    const PackArgs p = j.a;
    const int stage = (int)((long long)blockIdx.x - j.first_block);
    if ((long long)stage * 2048 >= j.total) return;
    const int KBc = (p.Cvalid + GEMM_BK - 1) / GEMM_BK, KB = p.ntaps * KBc;
    const int nb = stage / KB, kb = stage - nb * KB;
    const int tap = kb / KBc, cb = kb - tap * KBc;
    const float* wsrc = p.w + p.tap_idx[tap] * p.s_tap;
    uint8_t* stage_img = p.out + (size_t)stage * B_STAGE;
    // s_c == 1 (channels contiguous in the source): 8 lanes read one row's 64 channels and write one whole 128-byte image
    // row.  Otherwise rows are contiguous in the source: per pass the block takes 32 rows x 64 channels, lanes run along
    // the rows for the loads (whole 128-byte lines per channel) and the (hi, lo) chunks are transposed through shared
    // memory so that the stores cover whole image rows as well.
    const bool chunk_fast = p.s_c == 1;
    const int tid = threadIdx.x;
#pragma unroll 2
    for (int pass = 0; pass < 8; ++pass) {
        const int nl = chunk_fast ? pass * 32 + (tid >> 3) : pass * 32 + (tid & 31);
        const int chunk = chunk_fast ? (tid & 7) : (tid >> 5);
        const int n = nb * GEMM_BN + nl;
        const int c0 = cb * GEMM_BK + chunk * 8;
        float v[8];
        const float* src = wsrc + (long long)c0 * p.s_c + (long long)n * p.s_n;
        if (chunk_fast && n < p.Nvalid && c0 + 8 <= p.Cvalid && !((reinterpret_cast<uintptr_t>(src)) & 15)) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = c0 + e;
                v[e] = (n < p.Nvalid && c < p.Cvalid) ? __ldg(src + (long long)e * p.s_c) : 0.f;
            }
        }
        uint8_t* img = stage_img + (size_t)(nl >> 7) * B_SLOT;      // [CTA 0: hi | lo][CTA 1: hi | lo], 128 rows each
        if (chunk_fast) {
            const int nr = nl & 127;
            store_split(img, img + B_PLANE, nr * 128 + ((chunk ^ (nr & 7)) << 4), v);
        } else {
            uint4 vh, vl;
            split8(v, vh, vl);
            s_t[pass & 1][0][tid & 31][chunk] = vh; s_t[pass & 1][1][tid & 31][chunk] = vl;
            __syncthreads();                            // (two buffers: the next pass may start writing the other one)
            const int r2 = tid >> 3, ch2 = tid & 7;
            const int nr = ((pass * 32) & 127) | r2;
            const uint32_t off = nr * 128 + ((ch2 ^ (nr & 7)) << 4);
            *reinterpret_cast<uint4*>(img + off) = s_t[pass & 1][0][r2][ch2];
            *reinterpret_cast<uint4*>(img + B_PLANE + off) = s_t[pass & 1][1][r2][ch2];
        }
    }
}

void plan_job(PackJob* j, const float* w, int ntaps, const int* tap_idx, long long s_tap, long long s_c, long long s_n,
              int Cvalid, int Nvalid, void* out) {
    j->a.w = w; j->a.ntaps = ntaps;
    for (int i = 0; i < 3; ++i) j->a.tap_idx[i] = i < ntaps ? tap_idx[i] : 0;
    j->a.s_tap = s_tap; j->a.s_c = s_c; j->a.s_n = s_n; j->a.Cvalid = Cvalid; j->a.Nvalid = Nvalid;
    j->a.out = reinterpret_cast<uint8_t*>(out);
    j->total = (long long)(pack_image_bytes(ntaps, Cvalid, Nvalid) / B_STAGE) * 2048;
    j->first_block = 0;
}

// ------------------------------------------------------------------------------------------------ helpers
void conv_offsets(int k, int rate, int padding, int in_shift, int* off) {
    const int total = (k - 1) * rate;
    const int left = padding == OPH_PAD_CAUSAL ? total : total / 2;
    for (int j = 0; j < 3; ++j) off[j] = j < k ? j * rate - left - in_shift : 0;
}

int rows_grid(long long rows, int wpb) {
    long long g = (rows + wpb - 1) / wpb;
    const long long cap = 148 * 8;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

bool vec_ok(int C, long long ld0, long long ld1 = 4, long long ld2 = 4, long long ld3 = 4, long long ld4 = 4) {
    return (C == 256 || C == 512 || C == 1024) && !((ld0 | ld1 | ld2 | ld3 | ld4) & 3);
}
int bwd_grid(long long rows) {          // few, long-lived warps so that per-channel sums stay in registers
    long long g = (rows + 8 * 8 - 1) / (8 * 8);
    return (int)(g < 1 ? 1 : (g > 148 * 4 ? 148 * 4 : g));
}

int launch_ln_act_fwd(const float* z, long long ldz, const float* gamma, const float* beta, const oph_act* yo,
                      float* y_sig, long long ldys, float* stats, long long rows, int C, int act, int norm, float drop_p,
                      uint64_t seed, const long long* step, cudaStream_t st) {
    const int grid = rows_grid(rows, 8);
    float* y = yo->f32; const long long ldy = yo->ld;
    char pdesc[64] = ""; if (g_prof_on) snprintf(pdesc, sizeof(pdesc), "ln_fwd rows=%lld C=%d", (long long)rows, C);
    ProfScope ps(OPH_TAG_ROW_FWD, (double)rows * C * (yo->hi ? 12.0 : 8.0), st, pdesc);
    if (yo->hi && (!yo->lo || (yo->ldp & 7))) return fail(OPH_EINVAL, "split-bf16 output planes need 16-byte aligned rows%s");
    if (vec_ok(C, ldz, ldy, y_sig ? ldys : 4)) {
#define OPH_LAUNCH(V) launch_cfg(grid, 256, 0, st)(ln_act_fwd_vec_kernel<V>, z, ldz, gamma, beta, y, ldy, y_sig, ldys, yo->hi, yo->lo, yo->ldp, stats, (int)rows, act, norm, drop_p, seed, step)
        if (C == 256) OPH_LAUNCH(2); else if (C == 512) OPH_LAUNCH(4); else OPH_LAUNCH(8);
#undef OPH_LAUNCH
    } else if (!((ldz | ldy | (y_sig ? ldys : 4) | (yo->hi ? yo->ldp : 4)) & 3) && C <= 1280 && !(g_gemm_dbg_flags_host & 1048576)) {
        // other widths (80, 513, 1025): 16-byte accesses over the padded rows, elements >= C masked
        const int nv = (C + 3) / 4;
        const size_t sm = (2 * (size_t)((C + 3) & ~3) + 32) * sizeof(float);
#define OPH_LAUNCH(V, W) launch_cfg(rows_grid(rows, 8 / W), 256, sm, st)(ln_act_fwd_any_kernel<V, W>, z, ldz, gamma, beta, y, ldy, y_sig, ldys, yo->hi, yo->lo, yo->ldp, stats, (int)rows, C, act, norm, drop_p, seed, step)
        // (MAXV float4 per lane, WPR warps per row); (3, 2) / (3, 4) as in the backward kernel measured slower here: 226 vs
        // 178 us for 111 360 rows x 513 channels
        if (nv <= 32) OPH_LAUNCH(1, 1); else if (nv <= 160) OPH_LAUNCH(5, 1); else OPH_LAUNCH(5, 2);
#undef OPH_LAUNCH
    } else {
        launch_cfg(grid, 256, 0, st)(ln_act_fwd_kernel, z, ldz, gamma, beta, y, ldy, y_sig, ldys, yo->hi, yo->lo, yo->ldp, stats, (int)rows, C, act, norm, drop_p, seed, step);
    }
    return check_launch("ln_act_fwd_kernel");
}

// dz planes: when the vectorised kernel applies, dz is written ONLY as split-bf16 planes into the caller's scratch
// (hi plane [rows][C] followed by the lo plane: the same bytes as fp32 [rows][C]); *dzp tells the GEMMs what to read.
void dz_as_planes(float* dz, long long rows, int C, OperandMap* m) {
    m->hi = reinterpret_cast<const unsigned short*>(dz);
    m->lo = m->hi + rows * C;
    m->ld = C; m->ptr = nullptr;
}

int launch_ln_act_bwd(const float* dy, long long lddy, const float* z, long long ldz, const float* stats,
                      const float* gamma, const float* beta, float* dz, long long lddz, float* dgamma, float* dbeta,
                      float* dbias, long long rows, int C, int act, int norm, float drop_p, uint64_t seed,
                      const long long* step, OperandMap* dzmap, cudaStream_t st) {
    const size_t smem = 3 * (size_t)C * sizeof(float);
    dzmap->ptr = dz; dzmap->ld = lddz; dzmap->hi = dzmap->lo = nullptr;
    char pdesc[64] = ""; if (g_prof_on) snprintf(pdesc, sizeof(pdesc), "ln_bwd rows=%lld C=%d", (long long)rows, C);
    ProfScope ps(OPH_TAG_ROW_BWD, (double)rows * C * 12.0, st, pdesc);
    if (norm && vec_ok(C, lddy, ldz, lddz) && lddz >= C && !(g_gemm_dbg_flags_host & 32768)) {
        dz_as_planes(dz, rows, C, dzmap);
        unsigned short* h = const_cast<unsigned short*>(dzmap->hi); unsigned short* l = const_cast<unsigned short*>(dzmap->lo);
        const int wpr = C / 256, groups = 8 / wpr, depth = 3;
        long long gl = (rows + groups * 4 - 1) / (groups * 4);
        // blocks of 8 warps: 3 per SM (80 registers, ~55 KB of shared memory each); 2 per SM with flag 4194304.  Same time
        // per launch alone (25 us), but 6.48 vs 6.58 ms per training step: more, shorter-lived blocks beside the GEMMs
        const int capw = (g_gemm_dbg_flags_host & 4194304) ? 296 : 444;
        const int gridw = (int)(gl < 1 ? 1 : (gl > capw ? capw : gl));
        const size_t smw = (5 * (size_t)C + 32) * sizeof(float) + (size_t)8 * depth * LNB_SLOT;
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(ln_act_bwd_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            cudaFuncSetAttribute(ln_act_bwd_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            cudaFuncSetAttribute(ln_act_bwd_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            attr_done = true;
        }
#define OPH_LAUNCH(W) launch_cfg(gridw, 256, smw, st)(ln_act_bwd_wide_kernel<W>, dy, lddy, z, ldz, stats, gamma, beta, h, l, C, dgamma, dbeta, dbias, (int)rows, act, drop_p, seed, step, depth)
        if (C == 256) OPH_LAUNCH(1); else if (C == 512) OPH_LAUNCH(2); else OPH_LAUNCH(4);
#undef OPH_LAUNCH
    } else if (vec_ok(C, lddy, ldz, lddz) && C <= 512 && lddz >= C) {
        const int grid = bwd_grid(rows);
        dz_as_planes(dz, rows, C, dzmap);
        unsigned short* h = const_cast<unsigned short*>(dzmap->hi); unsigned short* l = const_cast<unsigned short*>(dzmap->lo);
        if (C == 256) launch_cfg(grid, 256, smem, st)(ln_act_bwd_vec_kernel<2>, dy, lddy, z, ldz, stats, gamma, beta, nullptr, 0, h, l, C, dgamma, dbeta, dbias, (int)rows, act, norm, drop_p, seed, step);
        else          launch_cfg(grid, 256, smem, st)(ln_act_bwd_vec_kernel<4>, dy, lddy, z, ldz, stats, gamma, beta, nullptr, 0, h, l, C, dgamma, dbeta, dbias, (int)rows, act, norm, drop_p, seed, step);
    } else {
        // other widths (80, 513, 1025, 1024): planes with rows padded to 8 elements when the caller's dz buffer holds them
        const long long ldp = (C + 7) / 8 * 8;
        unsigned short* h = nullptr; unsigned short* l = nullptr;
        if (lddz >= ldp && !(reinterpret_cast<uintptr_t>(dz) & 15)) {
            h = reinterpret_cast<unsigned short*>(dz); l = h + rows * ldp;
            dzmap->hi = h; dzmap->lo = l; dzmap->ld = ldp; dzmap->ptr = nullptr;
        }
        if (!((lddy | ldz | lddz) & 3) && C <= 1280 && !(g_gemm_dbg_flags_host & 1048576)) {
            const int nv = (C + 3) / 4;
            const size_t sm = (5 * (size_t)((C + 3) & ~3) + 32) * sizeof(float);      // gamma, beta, partial moments, 3 column sums
            const int wpr = nv <= 96 ? 1 : nv <= 192 ? 2 : 4;
            long long gl = (rows + (8 / wpr) * 8 - 1) / ((8 / wpr) * 8);                  // >= 8 rows per row group
            const int gridb = (int)(gl < 1 ? 1 : (gl > 148 * 3 ? 148 * 3 : gl));          // three resident blocks per SM
#define OPH_LAUNCH(V, W) launch_cfg(gridb, 256, sm, st)(ln_act_bwd_any_kernel<V, W>, dy, lddy, z, ldz, stats, gamma, beta, dz, lddz, h, l, ldp, dgamma, dbeta, dbias, (int)rows, C, act, norm, drop_p, seed, step)
            // 80 -> (1,1), up to 384 -> (3,1), 513 -> (3,2), 1025 -> (3,4): 36 accumulators per lane instead of 60 (80 registers,
            // three blocks per SM) and no pass in which a single lane works; 414 -> 338 us for 111 360 rows x 513 channels
            if (nv <= 32) OPH_LAUNCH(1, 1); else if (nv <= 96) OPH_LAUNCH(3, 1); else if (nv <= 192) OPH_LAUNCH(3, 2); else OPH_LAUNCH(3, 4);
#undef OPH_LAUNCH
        } else {
            launch_cfg(rows_grid(rows, 8), 256, smem, st)(ln_act_bwd_kernel, dy, lddy, z, ldz, stats, gamma, beta, dz, lddz, h, l, ldp, dgamma, dbeta, dbias, (int)rows, C, act, norm, drop_p, seed, step);
        }
    }
    return check_launch("ln_act_bwd_kernel");
}

// weight gradient: dW[tap][m][n] += sum_r Amap(r,tap)[m] * Bmap(r,tap)[n]   (operands: fp32 or split-bf16 planes)
int launch_wgrad(const OperandMap& a, int M, const int* a_off, int aL, int aLs, int a_mul,
                 const OperandMap& b, int N, const int* b_off, int bL, int bLs, int b_mul,
                 int R, int taps, float* dw, long long ldc, cudaStream_t st) {
    GemmArgs g = blank();
    g.a_mode = A_MNMAJOR; g.b_mode = B_MNMAJOR;
    g.A = a; g.A.L = aL; g.A.Ls = aLs; g.A.mul = a_mul;
    g.Bm = b; g.Bm.L = bL; g.Bm.Ls = bLs; g.Bm.mul = b_mul;
    for (int j = 0; j < 3; ++j) { g.A.off[j] = j < taps ? a_off[j] : 0; g.Bm.off[j] = j < taps ? b_off[j] : 0; }
    g.M = M; g.N = N; g.Kc = R; g.ytaps = taps; g.c_tap_stride = (long long)M * ldc;
    g.C = dw; g.ldc = ldc; g.atomic = 1; g.z_mode = Z_SPLITK; g.tag = OPH_TAG_WGRAD;
    // both operands pre-split and un-strided: feed them with TMA; the reduction then runs over (item, 64-step block) pairs
    g.r_tma = 0;
    const int items = aL > 0 ? R / aL : 0;
    if (g_use_tma && g.A.hi && g.Bm.hi && a_mul == 1 && b_mul == 1 && aL == aLs && bL == bLs && aL == bL && items * aL == R &&
        make_plane_tmap(&g.tmA_hi, g.A.hi, M, aL, items, g.A.ld, GEMM_BK) && make_plane_tmap(&g.tmA_lo, g.A.lo, M, aL, items, g.A.ld, GEMM_BK) &&
        make_plane_tmap(&g.tmB_hi, g.Bm.hi, N, bL, items, g.Bm.ld, GEMM_BK) && make_plane_tmap(&g.tmB_lo, g.Bm.lo, N, bL, items, g.Bm.ld, GEMM_BK)) {
        g.r_tma = 1; g.items = items;
    }
    // split the reduction so that the work units fill whole rounds of the 74 persistent CTA pairs (never 2 rounds + a sliver)
    const int base = cdiv(cdiv(M, GEMM_BM), 2) * cdiv(N, GEMM_BN) * taps;     // work units before split-K
    const int kblocks = g.r_tma ? items * cdiv(aL, GEMM_BK) : cdiv(R, GEMM_BK);
    const int max_splits = cdiv(kblocks, 4);
    // fewest split-K slices whose work units fill whole rounds of the 74 CTA pairs to >= 92 % (else the best found);
    // every slice costs one more RED pass over dW, so small counts win ties
    int splits = 1; double best_eff = 0.0;
    for (int sp = 1; sp <= max_splits && sp <= 64; ++sp) {
        const long long u = (long long)base * sp;
        const double eff = (double)u / (double)(cdiv((int)u, GEMM_MAX_PAIRS) * GEMM_MAX_PAIRS);
        if (eff > best_eff + 1e-9) { best_eff = eff; splits = sp; }
        if (eff >= 0.92 && u >= 2 * GEMM_MAX_PAIRS) break;
    }
    const int kb_chunk = cdiv(kblocks, splits);
    splits = cdiv(kblocks, kb_chunk);
    if (g.r_tma) { g.k_chunk = kb_chunk; g.Kc = kblocks; g.prof_k = R; }       // units of k-blocks
    else g.k_chunk = kb_chunk * GEMM_BK;
    return launch_gemm(g, splits, st);
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int oph_version(void) { return 100; }

// Host-side CRC-32C (Castagnoli, slicing-by-8) for the checkpoint writer / reader: a Text2Mel checkpoint with Adam slots
// is ~290 MB, which the pure-Python fallback of tf_checkpoint.py checksums at ~10 MB/s.
unsigned int oph_crc32c(const void* data, unsigned long long n, unsigned int crc) {
    static unsigned int T[8][256];
    static std::atomic<bool> ready{false};
    static std::mutex mu;
    if (!ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(mu);
        if (!ready.load(std::memory_order_relaxed)) {
            for (unsigned int i = 0; i < 256; ++i) {
                unsigned int c = i;
                for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
                T[0][i] = c;
            }
            for (int t = 1; t < 8; ++t)
                for (unsigned int i = 0; i < 256; ++i) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFF];
            ready.store(true, std::memory_order_release);
        }
    }
    const unsigned char* p = static_cast<const unsigned char*>(data);
    unsigned int c = crc ^ 0xFFFFFFFFu;
    while (n >= 8) {
        unsigned long long w;
        memcpy(&w, p, 8);                                   // (little-endian hosts only, like the rest of the format code)
        w ^= c;
        c = T[7][w & 0xFF] ^ T[6][(w >> 8) & 0xFF] ^ T[5][(w >> 16) & 0xFF] ^ T[4][(w >> 24) & 0xFF] ^
            T[3][(w >> 32) & 0xFF] ^ T[2][(w >> 40) & 0xFF] ^ T[1][(w >> 48) & 0xFF] ^ T[0][(w >> 56) & 0xFF];
        p += 8; n -= 8;
    }
    while (n--) c = T[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
const char* oph_last_error(void) { return g_err; }
long long oph_launch_count(void) { return g_launches.load(); }
int oph_gemm_debug_buffer(long long* dev_buf) { g_gemm_dbg = dev_buf; g_dbg_slots = 0; return OPH_OK; }
int oph_gemm_debug_ring(long long* dev_buf, int slots) { g_gemm_dbg = dev_buf; g_dbg_slots = dev_buf ? slots : 0; g_dbg_next = 0; g_dbg_desc.clear(); return OPH_OK; }
int oph_gemm_debug_ring_desc(int slot, char* out, int cap) {
    if (slot < 0 || slot >= (int)g_dbg_desc.size() || cap < 1) return OPH_EINVAL;
    snprintf(out, (size_t)cap, "%s", g_dbg_desc[slot].c_str());
    return OPH_OK;
}
int oph_gemm_debug_flags(int flags) { g_gemm_dbg_flags_host = flags; g_gemm_dbg_flags = flags & 0xA7; g_use_tma = !(flags & 8); g_use_split = !(flags & 16); g_use_pdl = (flags & 64) != 0; g_use_ln_fuse = (flags & 65536) != 0; g_use_hc_fuse = !(flags & 131072); g_use_attn_fuse = !(flags & 524288);
    g_hcb_depth = (flags >> 8) & 7;                    // 0 = automatic
    g_hcb_two = !(flags & 2048);
    if (g_hcb_depth == 1) g_hcb_depth = 2;
    if (g_hcb_depth > 4) g_hcb_depth = 4;
    return OPH_OK; }

int oph_wgrad_stream(oph_stream_t side, int enable) {
    g_wgrad_stream = S(side);
    g_wgrad_fork = enable != 0;
    return OPH_OK;
}

int oph_cache_config(int mode) {
    const cudaFuncCache m = mode == 1 ? cudaFuncCachePreferShared : mode == 2 ? cudaFuncCachePreferL1 : mode == 3 ? cudaFuncCachePreferEqual : cudaFuncCachePreferNone;
    if (cudaDeviceSetCacheConfig(m) != cudaSuccess) return check_launch("cudaDeviceSetCacheConfig");
    return OPH_OK;
}

// Self-test of the trap record: one block records a fake wait (warp 21 = the GEMM's copy-engine loader, barrier FULL_B[1])
// and traps.  The CUDA context is unusable afterwards: run it in a process of its own (tests/test_gpu_ops.py).
__global__ void trap_selftest_kernel() {
    if (threadIdx.x == 21 * 32) { oph_record_trap(229376u + 8u * (BAR_FULL_B + 1), 1u); __trap(); }
}
int oph_debug_trap_selftest(oph_stream_t stream) {
    ensure_trap_record();
    trap_selftest_kernel<<<1, GEMM_THREADS, 0, S(stream)>>>();
    cudaStreamSynchronize(S(stream));
    return check_launch("trap_selftest_kernel");
}

// The record a bounded barrier wait left before it trapped (0 words = none): out[0..3] raw, text = the decoded sentence.
int oph_last_trap(unsigned long long* out, char* text, int cap) {
    if (out) for (int i = 0; i < 4; ++i) out[i] = g_trap_host ? g_trap_host[i] : 0ull;
    if (text && cap > 0) { text[0] = 0; describe_trap(text, (size_t)cap); }
    return (g_trap_host && g_trap_host[0] == OPH_TRAP_MAGIC) ? 1 : 0;
}

int oph_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_prof_on = true;
    return OPH_OK;
}

// out[tag*3 + {0,1,2}] = launches, summed device milliseconds, summed algorithmic FLOPs (bytes for the row-wise tags)
int oph_profile_end(double* out) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = false;
    for (int i = 0; i < OPH_NUM_TAGS * 3; ++i) out[i] = 0.0;
    // OPH_PROF_DUMP=<path>: one line per launch (tag, ms, algorithmic work, shape) for tools/launch_table.py
    const char* dump = getenv("OPH_PROF_DUMP");
    FILE* df = dump && *dump ? fopen(dump, "a") : nullptr;
    if (df) fprintf(df, "# begin %zu launches\n", g_prof.size());
    for (auto& r : g_prof) {
        if (cudaEventSynchronize(r.e1) != cudaSuccess) return check_launch("profile_end");
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        const int t = (r.tag >= 0 && r.tag < OPH_NUM_TAGS) ? r.tag : 0;
        out[t * 3] += 1.0; out[t * 3 + 1] += ms; out[t * 3 + 2] += r.flops;
        if (df) fprintf(df, "%d %.3f %.6g %s\n", r.tag, ms * 1e3, r.flops, r.desc);
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    if (df) fclose(df);
    g_prof.clear();
    return OPH_OK;
}

size_t oph_conv_pack_bytes(int k, int Cin, int Cout, int deconv, int backward) {
    if (!deconv) return backward ? pack_image_bytes(k, Cout, Cin) : pack_image_bytes(k, Cin, Cout);
    if (backward) return pack_image_bytes(3, Cout, Cin);
    return pack_image_bytes(2, Cin, Cout) + pack_image_bytes(1, Cin, Cout);
}

int oph_conv_pack(const float* w, int k, int Cin, int Cout, int deconv, void* packed_fwd, void* packed_bwd,
                  oph_stream_t stream) {
    const int t012[3] = {0, 1, 2};
    if (!deconv) {
        if (k < 1 || k > 3) return fail(OPH_EINVAL, "conv_pack: k must be 1..3%s");
        const long long s_tap = (long long)Cin * Cout;
        if (packed_fwd) OPH_TRY(pack_image(w, k, t012, s_tap, Cout, 1, Cin, Cout, packed_fwd, S(stream)));
        if (packed_bwd) OPH_TRY(pack_image(w, k, t012, s_tap, 1, Cout, Cout, Cin, packed_bwd, S(stream)));
        return OPH_OK;
    }
    // [3][Cout][Cin]: forward B[n=co][c=ci]; even rows use taps (0,2), odd rows tap 1
    const long long s_tap = (long long)Cin * Cout;
    if (packed_fwd) {
        const int even[3] = {0, 2, 0}, odd[3] = {1, 0, 0};
        OPH_TRY(pack_image(w, 2, even, s_tap, 1, Cin, Cin, Cout, packed_fwd, S(stream)));
        uint8_t* second = reinterpret_cast<uint8_t*>(packed_fwd) + pack_image_bytes(2, Cin, Cout);
        OPH_TRY(pack_image(w, 1, odd, s_tap, 1, Cin, Cin, Cout, second, S(stream)));
    }
    if (packed_bwd) OPH_TRY(pack_image(w, 3, t012, s_tap, Cin, 1, Cout, Cin, packed_bwd, S(stream)));
    return OPH_OK;
}

size_t oph_pack_job_bytes(void) { return sizeof(PackJob); }

// Appends the (up to 3) jobs of one kernel to a HOST plan; *njobs and *nblocks are running totals.
int oph_pack_plan_add(void* plan_host, int capacity, int* njobs, long long* nblocks, const float* w, int k, int Cin,
                      int Cout, int deconv, void* packed_fwd, void* packed_bwd) {
    PackJob* plan = reinterpret_cast<PackJob*>(plan_host);
    const int t012[3] = {0, 1, 2};
    const long long s_tap = (long long)Cin * Cout;
    PackJob tmp[3];
    int n = 0;
    if (!deconv) {
        if (k < 1 || k > 3) return fail(OPH_EINVAL, "pack_plan_add: k must be 1..3%s");
        if (packed_fwd) plan_job(&tmp[n++], w, k, t012, s_tap, Cout, 1, Cin, Cout, packed_fwd);
        if (packed_bwd) plan_job(&tmp[n++], w, k, t012, s_tap, 1, Cout, Cout, Cin, packed_bwd);
    } else {
        if (packed_fwd) {
            const int even[3] = {0, 2, 0}, odd[3] = {1, 0, 0};
            plan_job(&tmp[n++], w, 2, even, s_tap, 1, Cin, Cin, Cout, packed_fwd);
            plan_job(&tmp[n++], w, 1, odd, s_tap, 1, Cin, Cin, Cout, reinterpret_cast<uint8_t*>(packed_fwd) + pack_image_bytes(2, Cin, Cout));
        }
        if (packed_bwd) plan_job(&tmp[n++], w, 3, t012, s_tap, Cin, 1, Cout, Cin, packed_bwd);
    }
    if (*njobs + n > capacity) return fail(OPH_EINVAL, "pack_plan_add: plan capacity exceeded%s");
    for (int i = 0; i < n; ++i) {
        tmp[i].first_block = *nblocks;
        *nblocks += tmp[i].total / 2048;                 // one block per stage
        plan[(*njobs)++] = tmp[i];
    }
    return OPH_OK;
}

int oph_pack_run(const void* plan_dev, int njobs, long long nblocks, oph_stream_t stream) {
    if (njobs <= 0 || nblocks <= 0) return OPH_OK;
    launch_cfg((unsigned)nblocks, 256, 0, S(stream))(pack_batch_kernel, reinterpret_cast<const PackJob*>(plan_dev), njobs);
    return check_launch("pack_batch_kernel");
}

// ------------------------------------------------------------------------------------------------ conv1d
static void set_operand(OperandMap& m, const oph_act* a) {
    m.ptr = a->f32; m.ld = a->ld; m.hi = m.lo = nullptr;
    if (a->hi) { m.hi = a->hi; m.lo = a->lo; m.ld = a->ldp; m.ptr = nullptr; }   // pre-split planes win: no conversion in the GEMM
}
static int check_act(const oph_act* a, const char* who) {
    if (!a || (!a->f32 && !a->hi)) return fail(OPH_EINVAL, "%s: activation has neither an fp32 view nor planes", who);
    if (a->hi && (!a->lo || (a->ldp & 7))) return fail(OPH_EINVAL, "%s: plane rows must be 16-byte aligned (ldp % 8 == 0)", who);
    return OPH_OK;
}

int oph_conv1d_fwd(const oph_act* x, const void* packed_w, const float* bias, const float* gamma, const float* beta,
                   float* z, long long ldz, float* stats, const oph_act* y, float* y_sig, long long ldys, int B, int L,
                   int Cin, int Cout, int k, int rate, int padding, int in_shift, int act, int norm, float drop_p,
                   uint64_t seed, const long long* step, oph_stream_t stream) {
    if (k < 1 || k > 3) return fail(OPH_EINVAL, "conv1d_fwd: k must be 1..3%s");
    OPH_TRY(check_act(x, "conv1d_fwd x"));
    GemmArgs g = blank();
    g.a_mode = A_KMAJOR; g.b_mode = B_PACKED;
    set_operand(g.A, x); g.A.L = L; g.A.Ls = L; g.A.mul = 1;
    conv_offsets(k, rate, padding, in_shift, g.A.off);
    g.Bpacked = packed_w; g.M = B * L; g.N = Cout; g.Kc = Cin; g.ntaps = k;
    g.C = z; g.ldc = ldz; g.bias = bias; g.tag = OPH_TAG_CONV_FWD;
    g.zero_bytes = (size_t)B * L * ldz * sizeof(float);
    if (norm && (!y->hi || (y->lo && !(y->ldp & 7)))) {         // ask for the LayerNorm / activation / dropout tail in the epilogue
        g.ln_gamma = gamma; g.ln_beta = beta; g.ln_stats = stats; g.ln_y = y->f32; g.ln_ldy = y->ld;
        g.ln_yhi = y->hi; g.ln_ylo = y->lo; g.ln_ldp = y->ldp; g.ln_ysig = y_sig; g.ln_ldys = ldys;
        g.ln_act = act; g.ln_drop = drop_p; g.ln_seed = seed; g.ln_step = step;
    }
    OPH_TRY(launch_gemm(g, 1, S(stream)));
    if (g.ln_fuse) return OPH_OK;
    return launch_ln_act_fwd(z, ldz, gamma, beta, y, y_sig, ldys, stats, (long long)B * L, Cout, act, norm, drop_p,
                             seed, step, S(stream));
}

int oph_conv1d_bwd(const float* dy, long long lddy, const oph_act* x, const float* z, long long ldz,
                   const float* stats, const void* packed_w_bwd, const float* gamma, const float* beta, float* dz,
                   long long lddz, float* dx, long long lddx, float* dw, float* dbias, float* dgamma, float* dbeta,
                   int B, int L, int Cin, int Cout, int k, int rate, int padding, int in_shift, int act, int norm,
                   float drop_p, uint64_t seed, const long long* step, oph_stream_t stream) {
    OPH_TRY(check_act(x, "conv1d_bwd x"));
    OperandMap dzm;
    OPH_TRY(launch_ln_act_bwd(dy, lddy, z, ldz, stats, gamma, beta, dz, lddz, dgamma, dbeta, dbias, (long long)B * L, Cout,
                              act, norm, drop_p, seed, step, &dzm, S(stream)));
    cudaStream_t ws = dw ? fork_wgrad(S(stream)) : S(stream);
    int off[3];
    conv_offsets(k, rate, padding, in_shift, off);
    if (dx) {
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_PACKED;
        g.A = dzm; g.A.L = L; g.A.Ls = L; g.A.mul = 1;
        for (int j = 0; j < 3; ++j) g.A.off[j] = -off[j];
        g.tag = OPH_TAG_DGRAD; g.Bpacked = packed_w_bwd; g.M = B * L; g.N = Cin; g.Kc = Cout; g.ntaps = k;
        g.C = dx; g.ldc = lddx;
        OPH_TRY(launch_gemm(g, 1, S(stream)));
    }
    if (dw) {
        const int zero[3] = {0, 0, 0};
        OperandMap xm; set_operand(xm, x);
        OPH_TRY(launch_wgrad(xm, Cin, off, L, L, 1, dzm, Cout, zero, L, L, 1, B * L, k, dw, Cout, ws));
    }
    return OPH_OK;
}

// ------------------------------------------------------------------------------------------------ normalize
// modules.normalize (modules.py:47-75) on its own: the conv tail without a conv in front of it.
int oph_normalize_fwd(const float* x, long long ldx, const float* gamma, const float* beta, const oph_act* y, float* stats,
                      long long rows, int C, oph_stream_t stream) {
    if (!x || !gamma || !beta || !y || !y->f32 || rows < 1 || C < 1) return fail(OPH_EINVAL, "normalize_fwd: missing operand%s");
    return launch_ln_act_fwd(x, ldx, gamma, beta, y, nullptr, 0, stats, rows, C, OPH_ACT_NONE, 1, 0.f, 0, nullptr, S(stream));
}
// dx (fp32) = dLN/dx . dy; dgamma / dbeta are accumulated into.  stats from the forward call.
int oph_normalize_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* stats,
                      const float* gamma, const float* beta, float* dx, long long lddx, float* dgamma, float* dbeta,
                      long long rows, int C, oph_stream_t stream) {
    if (!dy || !x || !stats || !gamma || !beta || !dx || !dgamma || !dbeta) return fail(OPH_EINVAL, "normalize_bwd: missing operand%s");
    launch_cfg(rows_grid(rows, 8), 256, 3 * (size_t)C * sizeof(float), S(stream))(ln_act_bwd_kernel, dy, lddy, x, ldx, stats, gamma, beta,
        dx, lddx, (unsigned short*)nullptr, (unsigned short*)nullptr, 0ll, dgamma, dbeta, (float*)nullptr, (int)rows, C, OPH_ACT_NONE, 1,
        0.f, 0ull, (const long long*)nullptr);
    return check_launch("ln_act_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------ highway conv
int oph_hc_fwd(const oph_act* x, const void* packed_w, const float* bias, const float* g1, const float* b1,
               const float* g2, const float* b2, float* z, long long ldz, float* stats, const oph_act* y, int B, int L,
               int C, int k, int rate, int padding, int norm, float drop_p, uint64_t seed, const long long* step,
               oph_stream_t stream) {
    if (k < 1 || k > 3) return fail(OPH_EINVAL, "hc_fwd: k must be 1..3%s");
    OPH_TRY(check_act(x, "hc_fwd x"));
    if (!x->f32) return fail(OPH_EINVAL, "hc_fwd: the highway residual needs the fp32 view of x%s");
    GemmArgs g = blank();
    g.a_mode = A_KMAJOR; g.b_mode = B_PACKED;
    set_operand(g.A, x); g.A.L = L; g.A.Ls = L; g.A.mul = 1;
    conv_offsets(k, rate, padding, 0, g.A.off);
    g.Bpacked = packed_w; g.M = B * L; g.N = 2 * C; g.Kc = C; g.ntaps = k;
    g.C = z; g.ldc = ldz; g.bias = bias; g.tag = OPH_TAG_HC_FWD;
    g.zero_bytes = (size_t)B * L * ldz * sizeof(float);        // z is ours to overwrite: remainder units may be K-split
    const long long all_rows = (long long)B * L;
    const LccGate lcc = take_lcc(L);                            // per-speaker gates on H2 (modules.py:200-201): generic tail kernel
    const bool vec = vec_ok(C, ldz, x->ld, y->ld) && !lcc.table;
    if (y->hi && (y->ldp & 7)) return fail(OPH_EINVAL, "hc_fwd: plane rows must be 16-byte aligned%s");
    if (y->hi && !vec && !lcc.table) return fail(OPH_EINVAL, "hc_fwd: output planes need C in {256,512,1024}%s");
    // Highway tail in the same launch (gemm_tc.cuh, hc_fused): row-tile pairs run as super-units whose epilogue warps finish
    // the rows.  A super-unit pins all column blocks of a row tile to one CTA pair, so the schedule is only as good as
    // (row tiles) / (74 pairs) rounds up: used when the last round is full or at least 70 % full.  Otherwise (e.g. 109 row
    // tiles at B=32, T=870: 74 + 35) the tail stays its own launch -- measured on B200, fusing the 74 and leaving the 35 to
    // plain units + a partial tail launch gives the same step time (7.02 vs 6.97 ms) and a slower conv launch.  The kernel
    // supports that mixed schedule (flag 262144 selects it) and the tests cover it.
    long long done_rows = 0;
    if (g_use_hc_fuse && vec && norm && !lcc.table && y->f32 && !(y->hi && (y->ldp & 3)) && !(g_gemm_dbg_flags_host & 4096)) {
        const int MP = cdiv(cdiv((int)all_rows, GEMM_BM), 2), P = GEMM_MAX_PAIRS;
        const int full = (MP / P) * P, rem = MP - full;
        const int F = (rem == 0 || rem * 10 >= P * 7) ? MP : ((g_gemm_dbg_flags_host & 262144) ? full : 0);
        if (F > 0) {
            g.hc_fused = F; g.hc_C = C;
            g.hc_x = x->f32; g.hc_ldx = x->ld; g.hc_g1 = g1; g.hc_b1 = b1; g.hc_g2 = g2; g.hc_b2 = b2;
            g.hc_y = y->f32; g.hc_ldy = y->ld; g.hc_yhi = y->hi; g.hc_ylo = y->lo; g.hc_ldp = y->ldp;
            g.hc_stats = stats; g.hc_drop = drop_p; g.hc_seed = seed; g.hc_step = step;
        }
    }
    OPH_TRY(launch_gemm(g, 1, S(stream)));
    if (g.hc_fused) done_rows = (long long)g.hc_fused * 2 * GEMM_BM < all_rows ? (long long)g.hc_fused * 2 * GEMM_BM : all_rows;
    if (done_rows >= all_rows) return OPH_OK;
    // the rows the GEMM launch did not finish: offset every row-indexed pointer
    const long long rows = all_rows - done_rows;
    z += done_rows * ldz;
    if (stats) stats += done_rows * 4;
    oph_act xr = *x, yr = *y;
    xr.f32 += done_rows * x->ld;
    yr.f32 += done_rows * y->ld;
    if (yr.hi) { yr.hi += done_rows * y->ldp; yr.lo += done_rows * y->ldp; }
    x = &xr; y = &yr;
    const unsigned long long row_base = (unsigned long long)done_rows;
    const int grid = rows_grid(rows, 8);
    char pdesc[64] = ""; if (g_prof_on) snprintf(pdesc, sizeof(pdesc), "hc_tail_fwd rows=%lld C=%d", (long long)rows, C);
    ProfScope ps(OPH_TAG_HC_ROW_FWD, (double)rows * C * (y->hi ? 20.0 : 16.0), S(stream), pdesc);
    if (vec && norm && !(g_gemm_dbg_flags_host & 4096)) {
        const int wpr = C / 256, groups = 8 / wpr;
        const int depth = ((g_gemm_dbg_flags_host & 8192) || C > 256) ? 3 : 2;         // ring slots per warp
        const int bps = depth == 3 ? 2 : 4;                                            // blocks of 8 warps per SM (shared memory)
        long long gl = (rows + groups * 2 - 1) / (groups * 2);
        const int gridw = (int)(gl < 1 ? 1 : (gl > 148 * bps ? 148 * bps : gl));
        const size_t smw = (4 * (size_t)C + 64) * sizeof(float) + (size_t)8 * depth * HCF_SLOT;
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(hc_post_fwd_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            cudaFuncSetAttribute(hc_post_fwd_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            cudaFuncSetAttribute(hc_post_fwd_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            attr_done = true;
        }
#define OPH_LAUNCH(W) launch_cfg(gridw, 256, smw, S(stream))(hc_post_fwd_wide_kernel<W>, z, ldz, x->f32, x->ld, g1, b1, g2, b2, y->f32, y->ld, y->hi, y->lo, y->ldp, stats, (int)rows, drop_p, seed, step, depth, (long long)row_base)
        if (C == 256) OPH_LAUNCH(1); else if (C == 512) OPH_LAUNCH(2); else OPH_LAUNCH(4);
#undef OPH_LAUNCH
    } else if (vec) {
#define OPH_LAUNCH(V) launch_cfg(grid, 256, 0, S(stream))(hc_post_fwd_vec_kernel<V>, z, ldz, x->f32, x->ld, g1, b1, g2, b2, y->f32, y->ld, y->hi, y->lo, y->ldp, stats, (int)rows, norm, drop_p, seed, step)
        if (C == 256) OPH_LAUNCH(2); else if (C == 512) OPH_LAUNCH(4); else OPH_LAUNCH(8);
#undef OPH_LAUNCH
    } else {
        launch_cfg(grid, 256, 0, S(stream))(hc_post_fwd_kernel, z, ldz, x->f32, x->ld, g1, b1, g2, b2, y->f32, y->ld, stats, (int)rows, C, norm, drop_p, seed, step,
                                                lcc, y->hi, y->lo, y->ldp);
    }
    return check_launch("hc_post_fwd_kernel");
}

int oph_hc_bwd(const float* dy, long long lddy, const oph_act* x, const float* z, long long ldz,
               const float* stats, const void* packed_w_bwd, const float* g1, const float* b1, const float* g2,
               const float* b2, float* dz, long long lddz, float* dxres, long long ldxr, float* dx, long long lddx,
               float* dw, float* dbias, float* dg1, float* db1, float* dg2, float* db2, int B, int L, int C, int k,
               int rate, int padding, int norm, float drop_p, uint64_t seed, const long long* step,
               oph_stream_t stream) {
    OPH_TRY(check_act(x, "hc_bwd x"));
    if (!x->f32) return fail(OPH_EINVAL, "hc_bwd: needs the fp32 view of x%s");
    const long long rows = (long long)B * L;
    const size_t smem = 6 * (size_t)C * sizeof(float);
    OperandMap dzm; dzm.ptr = dz; dzm.ld = lddz; dzm.hi = dzm.lo = nullptr;
    (void)dxres; (void)ldxr;
    const LccGate lcc = take_lcc(L);
    if (lcc.table && !lcc.scratch) return fail(OPH_EINVAL, "hc_bwd: the channel gates need a [B*L][C] scratch (oph_lcc_context)%s");
    {
    char pdesc[64] = ""; if (g_prof_on) snprintf(pdesc, sizeof(pdesc), "hc_tail_bwd rows=%lld C=%d", (long long)rows, C);
    ProfScope ps(OPH_TAG_ROW_BWD, (double)rows * C * 28.0, S(stream), pdesc);
    if (!lcc.table && norm && vec_ok(C, lddy, ldz, x->ld, lddz, lddx) && lddz >= 2 * C) {
        dz_as_planes(dz, rows, 2 * C, &dzm);
        unsigned short* h = const_cast<unsigned short*>(dzm.hi); unsigned short* l = const_cast<unsigned short*>(dzm.lo);
        const int wpr = C / 256, groups = 8 / wpr;
        long long gl = (rows + groups * 4 - 1) / (groups * 4);             // >= 4 rows per row group
        // two 8-warp blocks per SM; rows in flight per warp (ring slots): 3 where the shared memory of two blocks allows
        const int gmax = g_hcb_two ? 296 : 148;
        const int grid = (int)(gl < 1 ? 1 : (gl > gmax ? gmax : gl));
        const int depth = g_hcb_depth ? g_hcb_depth : (C == 256 ? 3 : 2);
        const size_t sm2 = (10 * (size_t)C + 64) * sizeof(float) + (size_t)8 * depth * HCB_SLOT;
        static bool attr_done = false;
        if (!attr_done) {
            const int mx = 200 * 1024;
            cudaFuncSetAttribute(hc_post_bwd_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            cudaFuncSetAttribute(hc_post_bwd_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            cudaFuncSetAttribute(hc_post_bwd_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            attr_done = true;
        }
#define OPH_LAUNCH(W) launch_cfg(grid, 256, sm2, S(stream))(hc_post_bwd_wide_kernel<W>, dy, lddy, z, ldz, x->f32, x->ld, stats, g1, b1, g2, b2, h, l, 2 * C, dx, lddx, dg1, db1, dg2, db2, dbias, (int)rows, drop_p, seed, step, depth)
        if (C == 256) OPH_LAUNCH(1); else if (C == 512) OPH_LAUNCH(2); else OPH_LAUNCH(4);
#undef OPH_LAUNCH
    } else {
        launch_cfg(rows_grid(rows, 8), 256, smem, S(stream))(hc_post_bwd_kernel, 
            dy, lddy, z, ldz, x->f32, x->ld, stats, g1, b1, g2, b2, dz, lddz, dx, lddx, dg1, db1, dg2, db2, dbias,
            (int)rows, C, norm, drop_p, seed, step, lcc);
    }
    }
    OPH_TRY(check_launch("hc_post_bwd_kernel"));
    cudaStream_t ws = dw ? fork_wgrad(S(stream)) : S(stream);
    int off[3];
    conv_offsets(k, rate, padding, 0, off);
    {
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_PACKED;
        g.A = dzm; g.A.L = L; g.A.Ls = L; g.A.mul = 1;
        for (int j = 0; j < 3; ++j) g.A.off[j] = -off[j];
        g.tag = OPH_TAG_DGRAD; g.Bpacked = packed_w_bwd; g.M = B * L; g.N = C; g.Kc = 2 * C; g.ntaps = k;
        // the row-wise kernel left do*(1-g) in dx; the GEMM accumulates W^T dz onto it with fire-and-forget RED.ADDs
        // (one add per element: deterministic), so the epilogue never waits on a global load
        g.C = dx; g.ldc = lddx; g.atomic = 1;
        OPH_TRY(launch_gemm(g, 1, S(stream)));
    }
    if (dw) {
        const int zero[3] = {0, 0, 0};
        OperandMap xm; set_operand(xm, x);
        OPH_TRY(launch_wgrad(xm, C, off, L, L, 1, dzm, 2 * C, zero, L, L, 1, B * L, k, dw, 2 * C, ws));
    }
    return OPH_OK;
}

// ------------------------------------------------------------------------------------------------ channel gates
int oph_lcc_context(const float* table, const int32_t* codes, float* scratch) {
    g_lcc = LccGate{table, codes, 0, scratch};
    return OPH_OK;
}
int oph_lcc_fwd(const float* y0, long long ld0, const float* table, const int32_t* codes, const oph_act* out, float* out_sig,
                long long lds, int B, int L, int C, oph_stream_t stream) {
    if (!y0 || !table || !codes || !out || !out->f32) return fail(OPH_EINVAL, "lcc_fwd: missing operand%s");
    const LccGate g = {table, codes, L, nullptr};
    launch_cfg(rows_grid((long long)B * L, 8), 256, 0, S(stream))(lcc_fwd_kernel, y0, ld0, g, out->f32, out->ld, out_sig, lds,
                                                                 out->hi, out->lo, out->ldp, (long long)B * L, C);
    return check_launch("lcc_fwd_kernel");
}
int oph_lcc_bwd(const float* dy, long long lddy, const float* y0, long long ld0, const float* table, const int32_t* codes,
                float* dy0, long long ldd0, float* scratch, int B, int L, int C, oph_stream_t stream) {
    if (!dy || !y0 || !table || !codes || !dy0 || !scratch) return fail(OPH_EINVAL, "lcc_bwd: missing operand%s");
    const LccGate g = {table, codes, L, scratch};
    launch_cfg(rows_grid((long long)B * L, 8), 256, 0, S(stream))(lcc_bwd_kernel, dy, lddy, y0, ld0, g, dy0, ldd0, (long long)B * L, C);
    return check_launch("lcc_bwd_kernel");
}
int oph_lcc_reduce(const float* scratch, const int32_t* codes, float* dtable, int B, int L, int C, oph_stream_t stream) {
    if (!scratch || !codes || !dtable) return fail(OPH_EINVAL, "lcc_reduce: missing operand%s");
    launch_cfg(dim3(cdiv(C, 32), B), 256, 0, S(stream))(lcc_reduce_kernel, scratch, codes, dtable, L, C);
    return check_launch("lcc_reduce_kernel");
}

// ------------------------------------------------------------------------------------------------ transposed conv
int oph_deconv_fwd(const oph_act* x, const void* packed_w, const float* bias, const float* gamma, const float* beta,
                   float* z, long long ldz, float* stats, const oph_act* y, int B, int L, int C, float drop_p,
                   uint64_t seed, const long long* step, oph_stream_t stream) {
    OPH_TRY(check_act(x, "deconv_fwd x"));
    for (int parity = 0; parity < 2; ++parity) {
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_PACKED;
        set_operand(g.A, x); g.A.L = L; g.A.Ls = L; g.A.mul = 1;
        g.A.off[0] = 0; g.A.off[1] = -1; g.A.off[2] = 0;       // even: W0.x[i] + W2.x[i-1]; odd: W1.x[i]
        g.ntaps = parity == 0 ? 2 : 1;
        g.Bpacked = reinterpret_cast<const uint8_t*>(packed_w) + (parity ? pack_image_bytes(2, C, C) : 0);
        g.M = B * L; g.N = C; g.Kc = C;
        g.C = z; g.ldc = ldz; g.c_mul = 2; g.c_off = parity; g.bias = bias; g.tag = OPH_TAG_CONV_FWD;
        OPH_TRY(launch_gemm(g, 1, S(stream)));
    }
    return launch_ln_act_fwd(z, ldz, gamma, beta, y, nullptr, 0, stats, 2LL * B * L, C, OPH_ACT_NONE, 1, drop_p, seed,
                             step, S(stream));
}

int oph_deconv_bwd(const float* dy, long long lddy, const oph_act* x, const float* z, long long ldz,
                   const float* stats, const void* packed_w_bwd, const float* gamma, const float* beta, float* dz,
                   long long lddz, float* dx, long long lddx, float* dw, float* dbias, float* dgamma, float* dbeta,
                   int B, int L, int C, float drop_p, uint64_t seed, const long long* step, oph_stream_t stream) {
    OPH_TRY(check_act(x, "deconv_bwd x"));
    OperandMap dzm;
    OPH_TRY(launch_ln_act_bwd(dy, lddy, z, ldz, stats, gamma, beta, dz, lddz, dgamma, dbeta, dbias, 2LL * B * L, C,
                              OPH_ACT_NONE, 1, drop_p, seed, step, &dzm, S(stream)));
    cudaStream_t ws = dw ? fork_wgrad(S(stream)) : S(stream);
    if (dx) {   // dx[i] = W0^T dz[2i] + W1^T dz[2i+1] + W2^T dz[2i+2]
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_PACKED;
        g.A = dzm; g.A.L = L; g.A.Ls = 2 * L; g.A.mul = 2;
        g.A.off[0] = 0; g.A.off[1] = 1; g.A.off[2] = 2;
        g.tag = OPH_TAG_DGRAD; g.Bpacked = packed_w_bwd; g.M = B * L; g.N = C; g.Kc = C; g.ntaps = 3;
        g.C = dx; g.ldc = lddx;
        OPH_TRY(launch_gemm(g, 1, S(stream)));
    }
    if (dw) {   // dW[j][co][ci]: j=0: dz[2i] x[i]; j=1: dz[2i+1] x[i]; j=2: dz[2i] x[i-1]
        const int a_off[3] = {0, 1, 0}, b_off[3] = {0, 0, -1};
        OperandMap xm; set_operand(xm, x);
        OPH_TRY(launch_wgrad(dzm, C, a_off, L, 2 * L, 2, xm, C, b_off, L, L, 1, B * L, 3, dw, C, ws));
    }
    return OPH_OK;
}

// ------------------------------------------------------------------------------------------------ embedding
int oph_embed_fwd(const int32_t* ids, const float* table, float* out, long long ldo, int rows, int E, oph_stream_t stream) {
    launch_cfg(rows_grid(rows, 8), 256, 0, S(stream))(embed_fwd_kernel, ids, table, out, ldo, rows, E);
    return check_launch("embed_fwd_kernel");
}
int oph_embed_bwd(const int32_t* ids, const float* dout, long long ldo, float* dtable, int rows, int E, oph_stream_t stream) {
    launch_cfg(rows_grid(rows, 8), 256, 0, S(stream))(embed_bwd_kernel, ids, dout, ldo, dtable, rows, E);
    return check_launch("embed_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------ attention
static bool planes_ready(const oph_act* a) { return a && a->hi && a->lo && !(a->ldp & 7) && !(reinterpret_cast<uintptr_t>(a->hi) & 15) && !(reinterpret_cast<uintptr_t>(a->lo) & 15); }
// operand of a batched product: planes [items][rows][ldp] (contiguous items) when `planes`, else the fp32 view
static void att_operand(OperandMap& m, const oph_act* a, bool planes, int rows) {
    m.L = m.Ls = rows; m.mul = 1;
    if (planes) { m.hi = a->hi; m.lo = a->lo; m.ld = a->ldp; m.ptr = nullptr; }
    else { m.hi = m.lo = nullptr; m.ptr = a->f32; m.ld = a->ld; }
}

int oph_split_planes(const float* x, long long ldx, long long rows, int C, unsigned short* hi, unsigned short* lo,
                     long long ldp, oph_stream_t stream) {
    launch_cfg(rows_grid(rows, 8), 256, 0, S(stream))(split_planes_kernel, x, ldx, rows, C, hi, lo, ldp);
    return check_launch("split_planes_kernel");
}

static int guide_tensor(GuideTensor& G, const oph_guide* guide, int N, int T) {
    G = GuideTensor{nullptr, 0, 0, 0, 0, 1.f, 0, nullptr, nullptr, 0.f};
    if (!guide) return OPH_OK;
    if ((guide->col_g == nullptr) != (guide->col_h == nullptr))
        return fail(OPH_EINVAL, "attention: col_g and col_h come together (oph_attention_extra_fwd)%s");
    G.col_g = guide->col_g; G.col_h = guide->col_h; G.c_aout = guide->c_aout;
    if (!guide->w) return OPH_OK;
    if (guide->Ng < 1 || guide->Tg < 1 || guide->ld < guide->Tg || guide->item_stride < (long long)guide->Ng * guide->ld)
        return fail(OPH_EINVAL, "attention: malformed guide tensor (Ng, Tg >= 1, ld >= Tg, item_stride >= Ng * ld)%s");
    // the MSE variant pads alignments and targets with zeros, so target mass outside the batch's [N, T] block would add a
    // constant the kernels do not see (architectures.py:271-280): forced-alignment targets must fit the batch
    if (guide->mse && (guide->Ng > N || guide->Tg > T))
        return fail(OPH_EINVAL, "attention: forced-alignment targets larger than the batch's [N, T] block%s");
    G.w = guide->w; G.item_stride = guide->item_stride; G.ld = guide->ld; G.Ng = guide->Ng; G.Tg = guide->Tg;
    G.pad = guide->pad; G.mse = guide->mse ? 1 : 0;
    return OPH_OK;
}

int oph_attention_fwd(const oph_act* Q, const oph_act* K, const oph_act* V, const oph_act* A, const oph_act* Ro,
                      float* align_t, int32_t* argmax, const int32_t* prev_max, int win, double* att_acc, int maxN,
                      int maxT, float g_, int B, int T, int N, int d, const oph_guide* guide, oph_stream_t stream) {
    if (!Q || !K || !V || !A || !Ro || !Ro->f32) return fail(OPH_EINVAL, "attention_fwd: missing operand%s");
    GuideTensor G;
    OPH_TRY(guide_tensor(G, guide, N, T));
    float* R = Ro->f32; const long long ldr = Ro->ld;
    const long long ldA = A->ld;
    if (A->f32 && ldA < N) return fail(OPH_EINVAL, "attention_fwd: ldA < N%s");
    // all operands as split-bf16 planes: every tile of both products arrives through the copy engines
    const bool fed = g_use_tma && planes_ready(Q) && planes_ready(K) && planes_ready(V) && planes_ready(A);
    if (!fed && (!Q->f32 || !K->f32 || !V->f32)) return fail(OPH_EINVAL, "attention_fwd: operands need fp32 views or planes%s");
    // One kernel for the whole block (attn_fused.cuh) when Q, K, V come as planes and the shape is the dc_tts one
    if (g_use_attn_fuse && g_use_tma && planes_ready(Q) && planes_ready(K) && planes_ready(V) && d == 256 && N >= 1 && N <= 256 &&
        !(A->hi && !planes_ready(A)) && !(Ro->hi && !planes_ready(Ro)) && !(ldr & 3)) {
        AttnArgs a;
        memset(&a, 0, sizeof(a));
        a.B = B; a.T = T; a.N = N; a.d = d;
        a.Npad = cdiv(N, 16) * 16; a.nblk = cdiv(a.Npad, 64); a.vslots = a.nblk <= 3 ? 2 : 1;
        a.scale = 1.0f / sqrtf((float)d);
        a.prev_max = prev_max; a.win = win;
        a.A = A->f32; a.ldA = ldA; a.Ahi = A->hi; a.Alo = A->lo; a.ldAp = A->ldp;
        a.align_t = align_t; a.argmax = argmax; a.att_acc = att_acc; a.maxN = maxN; a.maxT = maxT; a.g = g_; a.G = G;
        a.R = R; a.ldr = ldr; a.Rhi = Ro->hi; a.Rlo = Ro->lo; a.ldrp = Ro->ldp;
        if (make_plane_tmap(&a.tmQ_hi, Q->hi, d, T, B, Q->ldp, ATT_BM) && make_plane_tmap(&a.tmQ_lo, Q->lo, d, T, B, Q->ldp, ATT_BM) &&
            make_plane_tmap(&a.tmK_hi, K->hi, d, N, B, K->ldp, a.Npad) && make_plane_tmap(&a.tmK_lo, K->lo, d, N, B, K->ldp, a.Npad) &&
            make_plane_tmap(&a.tmV_hi, V->hi, d, N, B, V->ldp, 64) && make_plane_tmap(&a.tmV_lo, V->lo, d, N, B, V->ldp, 64)) {
            static bool attr_done = false;
            if (!attr_done) {
                if (cudaFuncSetAttribute(attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) != cudaSuccess)
                    return check_launch("cudaFuncSetAttribute(attn_fused)");
                attr_done = true;
            }
            const int units = B * cdiv(T, ATT_BM);
            ProfScope ps(OPH_TAG_ATTENTION, 4.0 * B * T * (double)N * d, S(stream));
            ensure_trap_record(S(stream));
            launch_cfg(units < 148 ? units : 148, ATT_THREADS, ATT_SMEM, S(stream))(attn_fused_kernel, a);
            return check_launch("attn_fused_kernel");
        }
    }
    if (!A->f32) return fail(OPH_EINVAL, "attention_fwd: this shape runs as three launches and needs A (fp32 [B][T][ldA]) as scratch%s");
    {   // S = Q K^T / sqrt(d)
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_KMAJOR;
        att_operand(g.A, Q, fed, T); att_operand(g.Bm, K, fed, N);
        g.M = T; g.N = N; g.Kc = d;
        g.tag = OPH_TAG_ATTENTION; g.z_mode = Z_BATCH; g.a_zs = (long long)T * Q->ld; g.b_zs = (long long)N * K->ld; g.c_zs = (long long)T * ldA;
        g.C = A->f32; g.ldc = ldA; g.alpha = 1.0f / sqrtf((float)d);
        OPH_TRY(launch_gemm(g, B, S(stream)));
    }
    launch_cfg(rows_grid((long long)B * T, 8), 256, 0, S(stream))(softmax_fwd_kernel, A->f32, ldA, B, T, N, prev_max, win, align_t,
                                                                 argmax, att_acc, maxN, maxT, g_, G, fed ? A->hi : nullptr,
                                                                 fed ? A->lo : nullptr, A->ldp);
    OPH_TRY(check_launch("softmax_fwd_kernel"));
    {   // R = A V
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_MNMAJOR;
        att_operand(g.A, A, fed, T); att_operand(g.Bm, V, fed, N);
        g.M = T; g.N = d; g.Kc = N;
        g.tag = OPH_TAG_ATTENTION; g.z_mode = Z_BATCH; g.a_zs = (long long)T * ldA; g.b_zs = (long long)N * V->ld; g.c_zs = (long long)T * ldr;
        g.C = R; g.ldc = ldr;
        if (planes_ready(Ro) && !(ldr & 3)) { g.Chi = Ro->hi; g.Clo = Ro->lo; g.ldcp = Ro->ldp; g.cp_zs = (long long)T * Ro->ldp; }
        OPH_TRY(launch_gemm(g, B, S(stream)));
    }
    return OPH_OK;
}

int oph_attention_guide_sum(const float* A, long long ldA, int B, int T, int N, double* att_acc, int maxN, int maxT, float g_,
                            const oph_guide* guide, oph_stream_t stream) {
    if (!A || !att_acc || B < 1 || T < 1 || N < 1 || ldA < N) return fail(OPH_EINVAL, "attention_guide_sum: missing operand%s");
    GuideTensor G;
    OPH_TRY(guide_tensor(G, guide, N, T));
    launch_cfg(rows_grid((long long)B * T, 8), 256, 0, S(stream))(guide_sum_kernel, A, ldA, B, T, N, att_acc, maxN, maxT, g_, G);
    return check_launch("guide_sum_kernel");
}

int oph_attention_extra_fwd(const float* A, long long ldA, int B, int T, int N, float c_cdp, float c_ain, float* col_g,
                            float* col_h, double* acc3, oph_stream_t stream) {
    if (!A || !col_g || !col_h || !acc3 || B < 1 || T < 1 || N < 1) return fail(OPH_EINVAL, "attention_extra_fwd: missing operand%s");
    launch_cfg(dim3(cdiv(N, 32), B), 256, 0, S(stream))(att_extra_kernel, A, ldA, B, T, N, c_cdp, c_ain, col_g, col_h, acc3);
    return check_launch("att_extra_kernel");
}

int oph_attention_extra_finalize(const double* acc3, float* comps8, int B, int T, int N, float w_cdp, float w_ain,
                                 float w_aout, int add_to_total, oph_stream_t stream) {
    if (T < 2 || N < 2) return fail(OPH_EINVAL, "attention_extra_finalize: the entropies are normalised by log T and log N%s");
    launch_cfg(1, 1, 0, S(stream))(att_extra_finalize_kernel, acc3, comps8, (double)B * N, (double)B * T, log((double)T),
                                   log((double)N), w_cdp, w_ain, w_aout, add_to_total);
    return check_launch("att_extra_finalize_kernel");
}

int oph_attention_bwd(const oph_act* dR, const oph_act* Q, const oph_act* K, const oph_act* V, const oph_act* A,
                      const oph_act* dA, float* dQ, long long lddq, const float* dq_addend, long long ldqa, float* dK,
                      long long lddk, float* dV, long long lddv, float att_coef, int maxN, int maxT, float g_, int B,
                      int T, int N, int d, const oph_guide* guide, oph_stream_t stream) {
    if (!dR || !Q || !K || !V || !A || !dA || !A->f32 || !dA->f32) return fail(OPH_EINVAL, "attention_bwd: missing operand%s");
    GuideTensor G;
    OPH_TRY(guide_tensor(G, guide, N, T));
    const float scale = 1.0f / sqrtf((float)d);
    const long long ldA = A->ld;
    if (dA->ld != ldA) return fail(OPH_EINVAL, "attention_bwd: dA must share A's row stride%s");
    const bool fed = g_use_tma && planes_ready(dR) && planes_ready(Q) && planes_ready(K) && planes_ready(V) &&
                     planes_ready(A) && planes_ready(dA);
    if (!fed && (!dR->f32 || !Q->f32 || !K->f32 || !V->f32)) return fail(OPH_EINVAL, "attention_bwd: operands need fp32 views or planes%s");
    {   // dV[n][:] = sum_t A[t][n] dR[t][:]
        GemmArgs g = blank();
        g.a_mode = A_MNMAJOR; g.b_mode = B_MNMAJOR;
        att_operand(g.A, A, fed, T); att_operand(g.Bm, dR, fed, T);
        g.M = N; g.N = d; g.Kc = T;
        g.tag = OPH_TAG_ATTENTION; g.z_mode = Z_BATCH; g.a_zs = (long long)T * ldA; g.b_zs = (long long)T * dR->ld; g.c_zs = (long long)N * lddv;
        g.C = dV; g.ldc = lddv;
        OPH_TRY(launch_gemm(g, B, S(stream)));
    }
    {   // dA = dR V^T
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_KMAJOR;
        att_operand(g.A, dR, fed, T); att_operand(g.Bm, V, fed, N);
        g.M = T; g.N = N; g.Kc = d;
        g.tag = OPH_TAG_ATTENTION; g.z_mode = Z_BATCH; g.a_zs = (long long)T * dR->ld; g.b_zs = (long long)N * V->ld; g.c_zs = (long long)T * ldA;
        g.C = dA->f32; g.ldc = ldA;
        OPH_TRY(launch_gemm(g, B, S(stream)));
    }
    launch_cfg(rows_grid((long long)B * T, 8), 256, 0, S(stream))(softmax_bwd_kernel, A->f32, ldA, dA->f32, ldA, B, T, N, att_coef, maxN,
                                                                 maxT, g_, G, fed ? dA->hi : nullptr, fed ? dA->lo : nullptr, dA->ldp);
    OPH_TRY(check_launch("softmax_bwd_kernel"));
    {   // dQ = dS K / sqrt(d) (+ direct path)
        GemmArgs g = blank();
        g.a_mode = A_KMAJOR; g.b_mode = B_MNMAJOR;
        att_operand(g.A, dA, fed, T); att_operand(g.Bm, K, fed, N);
        g.M = T; g.N = d; g.Kc = N;
        g.tag = OPH_TAG_ATTENTION; g.z_mode = Z_BATCH; g.a_zs = (long long)T * ldA; g.b_zs = (long long)N * K->ld; g.c_zs = (long long)T * lddq;
        g.C = dQ; g.ldc = lddq; g.alpha = scale;
        if (dq_addend) {
            if ((long long)T * ldqa != g.c_zs || ldqa != lddq) return fail(OPH_EINVAL, "attention_bwd: dq_addend must share dQ's layout%s");
            g.addend = dq_addend; g.ld_add = ldqa;
        }
        OPH_TRY(launch_gemm(g, B, S(stream)));
    }
    {   // dK[n][:] = sum_t dS[t][n] Q[t][:] / sqrt(d)
        GemmArgs g = blank();
        g.a_mode = A_MNMAJOR; g.b_mode = B_MNMAJOR;
        att_operand(g.A, dA, fed, T); att_operand(g.Bm, Q, fed, T);
        g.M = N; g.N = d; g.Kc = T;
        g.tag = OPH_TAG_ATTENTION; g.z_mode = Z_BATCH; g.a_zs = (long long)T * ldA; g.b_zs = (long long)T * Q->ld; g.c_zs = (long long)N * lddk;
        g.C = dK; g.ldc = lddk; g.alpha = scale;
        OPH_TRY(launch_gemm(g, B, S(stream)));
    }
    return OPH_OK;
}

// ------------------------------------------------------------------------------------------------ losses / optimiser
int oph_recon_loss(const float* logits, long long ldl, const float* target, long long ldt, float* dlogits,
                   long long ldd, long long rows, int C, int squash, float w_l1, float w_bd, float w_l2,
                   double* acc, oph_stream_t stream) {
    launch_cfg(rows_grid(rows, 8), 256, 0, S(stream))(recon_loss_kernel, logits, ldl, target, ldt, dlogits, ldd, rows, C, squash,
                                                                w_l1, w_bd, w_l2, acc);
    return check_launch("recon_loss_kernel");
}
int oph_loss_finalize(const double* acc, float* out, double n_recon, double n_att, float w_l1, float w_bd,
                      float w_att, float w_l2, int has_att, int squash, oph_stream_t stream) {
    launch_cfg(1, 1, 0, S(stream))(loss_finalize_kernel, acc, out, n_recon, n_att, w_l1, w_bd, w_att, w_l2, has_att, squash);
    return check_launch("loss_finalize_kernel");
}
int oph_adam_prepare(const long long* global_step, float* lr_t, float lr0, float beta1, float beta2, int decay_lr,
                     float warmup, oph_stream_t stream) {
    launch_cfg(1, 1, 0, S(stream))(adam_prepare_kernel, global_step, lr_t, lr0, beta1, beta2, decay_lr, warmup);
    return check_launch("adam_prepare_kernel");
}
int oph_adam_clip(float* p, float* m, float* v, const float* g, long long n, const float* lr_t, float beta1,
                  float beta2, float eps, float clip, float grad_scale, oph_stream_t stream) {
    launch_cfg(148 * 4, 256, 0, S(stream))(adam_clip_kernel, p, m, v, g, n, lr_t, beta1, beta2, eps, clip, grad_scale);
    return check_launch("adam_clip_kernel");
}
int oph_step_inc(long long* global_step, oph_stream_t stream) {
    launch_cfg(1, 1, 0, S(stream))(step_inc_kernel, global_step);
    return check_launch("step_inc_kernel");
}

// ------------------------------------------------------------------------------------------------ incremental frame step
#define OPH_AR_SCRATCH_FLOATS (32ll * AR_MAXB * 1024)
size_t oph_ar_scratch_floats(void) { return (size_t)OPH_AR_SCRATCH_FLOATS; }

// slices of the reduction over k * Cin: ~128 CTAs per layer, at most 256 and at least 16 products per slice
static int ar_gemv(const float* x, long long x_item, long long ldx, const float* w, float* partial, int B, int Cin, int O,
                   int k, int rate, int in_shift, const int* frame, int* KS_out, cudaStream_t st) {
    if (B < 1 || B > AR_MAXB) return fail(OPH_EINVAL, "ar step: 1 <= B <= 16%s");
    if (k < 1 || k > 3 || Cin < 1 || O < 1 || O > 2048) return fail(OPH_EINVAL, "ar step: unsupported layer shape%s");
    const int K = k * Cin, tiles = cdiv(O, AR_COLS);
    int KS = cdiv(128, tiles);
    if (KS > cdiv(K, 16)) KS = cdiv(K, 16);
    if (KS < cdiv(K, AR_KSLICE_MAX)) KS = cdiv(K, AR_KSLICE_MAX);
    if (KS > 32) KS = 32;
    int kslice = cdiv(cdiv(K, KS), 4) * 4;
    if (kslice > AR_KSLICE_MAX) return fail(OPH_EINVAL, "ar step: k * Cin too large%s");
    KS = cdiv(K, kslice);
    if ((long long)KS * B * O > OPH_AR_SCRATCH_FLOATS) return fail(OPH_EINVAL, "ar step: scratch too small%s");
    launch_cfg(dim3(tiles, KS), 256, 0, st)(ar_gemv_kernel, w, x, x_item, ldx, Cin, O, k, rate, in_shift, frame, B, partial, kslice);
    *KS_out = KS;
    return check_launch("ar_gemv_kernel");
}

int oph_ar_conv_step(const float* x, long long x_item, long long ldx, const float* w, const float* bias,
                     const float* gamma, const float* beta, float* y, long long y_item, long long ldy, float* y_sig,
                     long long s_item, long long lds, float* scratch, int B, int Cin, int Cout, int k, int rate,
                     int in_shift, int act, const int* frame, oph_stream_t stream) {
    int KS = 0;
    OPH_TRY(ar_gemv(x, x_item, ldx, w, scratch, B, Cin, Cout, k, rate, in_shift, frame, &KS, S(stream)));
    launch_cfg(B, 256, (size_t)Cout * sizeof(float), S(stream))(ar_tail_kernel, scratch, KS, bias, Cout, 0, gamma, beta,
        (const float*)nullptr, (const float*)nullptr, act, (const float*)nullptr, 0ll, 0ll, y, y_item, ldy, y_sig, s_item, lds, frame, B);
    return check_launch("ar_tail_kernel");
}

int oph_ar_hc_step(const float* x, long long x_item, long long ldx, const float* w, const float* bias, const float* g1,
                   const float* b1, const float* g2, const float* b2, float* y, long long y_item, long long ldy,
                   float* scratch, int B, int C, int k, int rate, const int* frame, oph_stream_t stream) {
    int KS = 0;
    if ((g1 == nullptr) != (g2 == nullptr)) return fail(OPH_EINVAL, "ar_hc_step: H1 and H2 are normalised together%s");
    OPH_TRY(ar_gemv(x, x_item, ldx, w, scratch, B, C, 2 * C, k, rate, 0, frame, &KS, S(stream)));
    launch_cfg(B, 256, (size_t)2 * C * sizeof(float), S(stream))(ar_tail_kernel, scratch, KS, bias, 2 * C, 1, g1, b1, g2, b2, 0,
        x, x_item, ldx, y, y_item, ldy, (float*)nullptr, 0ll, 0ll, frame, B);
    return check_launch("ar_tail_kernel");
}

int oph_ar_encoder_step(const oph_ar_layer* layers, int nlayers, int B, const int* frame, oph_stream_t stream) {
    if (!layers || nlayers < 1 || nlayers > AR_ENC_MAX_LAYERS || B < 1) return fail(OPH_EINVAL, "ar_encoder_step: 1..16 layers%s");
    ArEncArgs a;
    memset(&a, 0, sizeof(a));
    a.nlayers = nlayers;
    for (int i = 0; i < nlayers; ++i) {
        const oph_ar_layer& s = layers[i];
        const int cols = (s.kind ? 2 : 1) * (s.C / AR_ENC_CLUSTER);
        if (s.C % AR_ENC_CLUSTER || (cols != 32 && cols != 64 && cols != 128) || s.k < 1 || s.k > 3 || s.Cin < 1 ||
            s.k * s.Cin > AR_ENC_MAXK || (s.kind && s.Cin != s.C) || !s.w || !s.x || !s.y ||
            ((s.g1 == nullptr) != (s.b1 == nullptr)) || (s.kind && ((s.g1 == nullptr) != (s.g2 == nullptr))))
            return fail(OPH_EINVAL, "ar_encoder_step: unsupported layer shape (use the per-layer calls)%s");
        ArEncLayer& d = a.L[i];
        d.w = s.w; d.bias = s.bias; d.g1 = s.g1; d.b1 = s.b1; d.g2 = s.g2; d.b2 = s.b2; d.x = s.x; d.y = s.y;
        d.x_item = s.x_item; d.ldx = s.ldx; d.y_item = s.y_item; d.ldy = s.ldy;
        d.Cin = s.Cin; d.C = s.C; d.k = s.k; d.rate = s.rate; d.kind = s.kind; d.act = s.act; d.in_shift = s.in_shift;
    }
    double wbytes = 0.0;
    for (int i = 0; i < nlayers; ++i) wbytes += 4.0 * layers[i].k * layers[i].Cin * layers[i].C * (layers[i].kind ? 2 : 1);
    ProfScope ps(OPH_TAG_AR_ENC, wbytes, S(stream));
    launch_cfg(dim3(AR_ENC_CLUSTER, B), AR_ENC_THREADS, 0, S(stream))(ar_encoder_kernel, a, frame);
    return check_launch("ar_encoder_kernel");
}

int oph_ar_window_gather(const float* Q, long long q_item, long long ldq, float* Qw, long long w_item, long long ldw,
                         int B, int d, int T, int W, int reach, const int* frame, oph_stream_t stream) {
    if (B < 1 || W < 1 || W > T || reach < 0) return fail(OPH_EINVAL, "ar_window_gather: need 1 <= W <= T%s");
    launch_cfg(dim3(W, B), 128, 0, S(stream))(ar_window_gather_kernel, Q, q_item, ldq, Qw, w_item, ldw, d, T, W, reach, frame);
    return check_launch("ar_window_gather_kernel");
}

int oph_ar_window_scatter(const float* Yw, long long yw_item, long long ldyw, float* Y, long long y_item, long long ldy,
                          int n_mels, const float* align_w, float* align_t, const int32_t* argmax_w, int32_t* prev,
                          int32_t* history, int B, int N, int T, int W, int reach, const int* frame,
                          oph_stream_t stream) {
    if (B < 1 || W < 1 || W > T || reach < 0) return fail(OPH_EINVAL, "ar_window_scatter: need 1 <= W <= T%s");
    launch_cfg(B, 256, 0, S(stream))(ar_window_scatter_kernel, Yw, yw_item, ldyw, Y, y_item, ldy, n_mels, align_w, align_t,
                                     argmax_w, prev, history, B, N, T, W, reach, frame);
    return check_launch("ar_window_scatter_kernel");
}

int oph_ar_advance(int32_t* frame, oph_stream_t stream) {
    launch_cfg(1, 1, 0, S(stream))(ar_advance_kernel, frame);
    return check_launch("ar_advance_kernel");
}

// ------------------------------------------------------------------------------------------------ raw GEMM (tests)
int oph_gemm_nt(const float* A, long long lda, const float* Bm, long long ldb, float* C, long long ldc,
                const float* bias, int M, int N, int K, int b_mode, float alpha, int batch, long long a_bs,
                long long b_bs, long long c_bs, oph_stream_t stream) {
    GemmArgs g = blank();
    g.a_mode = A_KMAJOR; g.b_mode = b_mode == 2 ? B_MNMAJOR : B_KMAJOR;
    g.A.ptr = A; g.A.ld = lda; g.A.L = M; g.A.Ls = M;
    g.Bm.ptr = Bm; g.Bm.ld = ldb; g.Bm.L = K; g.Bm.Ls = K;
    g.M = M; g.N = N; g.Kc = K; g.C = C; g.ldc = ldc; g.bias = bias; g.alpha = alpha;
    if (batch > 1) { g.tag = OPH_TAG_ATTENTION; g.z_mode = Z_BATCH; g.a_zs = a_bs; g.b_zs = b_bs; g.c_zs = c_bs; }
    return launch_gemm(g, batch, S(stream));
}
int oph_gemm_tn(const float* A, long long lda, const float* Bm, long long ldb, float* C, long long ldc, int M,
                int N, int R, int splits, oph_stream_t stream) {
    GemmArgs g = blank();
    g.a_mode = A_MNMAJOR; g.b_mode = B_MNMAJOR;
    g.A.ptr = A; g.A.ld = lda; g.A.L = R; g.A.Ls = R;
    g.Bm.ptr = Bm; g.Bm.ld = ldb; g.Bm.L = R; g.Bm.Ls = R;
    g.M = M; g.N = N; g.Kc = R; g.C = C; g.ldc = ldc; g.atomic = 1; g.z_mode = Z_SPLITK;
    if (splits < 1) splits = 1;
    g.k_chunk = cdiv(cdiv(R, splits), GEMM_BK) * GEMM_BK;
    return launch_gemm(g, cdiv(R, g.k_chunk), S(stream));
}

}  // extern "C"
