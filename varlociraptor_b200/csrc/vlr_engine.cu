// vlr_engine.cu — sm_100a kernel and C-ABI (include/vlr_engine.h) of the per-locus posterior engine.
//
// Kernel: persistent grid (resident CTAs per SM x 148 SMs), one warp per locus, loci handed out through a global
// atomic ticket so that the very uneven per-locus work (a few hundred to >100k per-read evaluations)
// balances dynamically. All per-locus logic lives in engine_core.cuh.
//
// Host: vlr_call_batch() streams a host batch through 6 slots (chunk of loci -> H2D -> kernel -> D2H, one CUDA
// stream per slot) so copies overlap compute; vlr_call_batch_device() launches on device-resident buffers.
// There is no CPU fallback: without a usable CUDA device vlr_ctx_create() fails with VLR_ERR_NO_DEVICE.
#include <cuda_fp16.h>
#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#ifdef __linux__
#include <sys/syscall.h>
#include <unistd.h>
#endif

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

// The engine core is instantiated twice by inclusion: a small variant (<= 3 samples, <= 8 events, tree depth <= 6:
// tumor-normal, trios) whose per-warp state is ~2 KB of shared memory, and the full-capacity variant.
#define VLR_VARIANT vlr_small
#define VLR_VAR_MAXS 3
#define VLR_VAR_MAXE 8
#define VLR_VAR_MAXD 6
#define VLR_VAR_WAVE 1 // + the wavefront pipeline (engine_wave.cuh) for two-level chain scenarios (tumor-normal)
#include "engine_core.cuh"
#define VLR_VARIANT vlr_full
#define VLR_VAR_MAXS VLR_MAX_SAMPLES
#define VLR_VAR_MAXE VLR_MAX_EVENTS
#define VLR_VAR_MAXD VLR_MAX_TREE_DEPTH
#include "engine_core.cuh"
#include "scenario_prep.h"
#include "contamination.cuh"

using namespace vlrcore;

namespace {

constexpr int THREADS = WARPS_PER_CTA * LANES; // WARPS_PER_CTA counts logical warps (engine_types.cuh)
constexpr int NBUF_MAX = 8; // chunk slots of vlr_call_batch (vlr_ctx::nbuf of them are used)

struct KernelParams {
    DevScenario sc;
    DevBatch b;
    DevResults r;
    WarpWs* ws;
    double* coef;
    double* be;
    unsigned long long* ticket;
    const int* locus_list;          // optional: the loci to process (deferred by the wavefront pipeline) ...
    const unsigned* locus_list_n;   // ... and their number (device counter)
    int coef_cap;      // reads per warp in the global arena
    int sm_reads;      // reads per warp in the shared-memory arena
    int ctx_stride;    // bytes of shared memory per warp for the Ctx
    int64_t be_stride; // doubles per warp
};


// Shared memory per CTA: [WARPS_PER_CTA x Ctx (uniform per-warp state)] [WARPS_PER_CTA x coefficient arena].
#define VLR_DEFINE_KERNEL(NS)                                                                                     \
    __global__ void __launch_bounds__(THREADS, VLR_MIN_CTAS) vlr_call_kernel_##NS(const __grid_constant__ KernelParams p) {  \
                const int w = group_in_cta();                                                                             \
        const int gw = blockIdx.x * WARPS_PER_CTA + w;                                                            \
        NS::Ctx& c = *reinterpret_cast<NS::Ctx*>(vlr_smem + (size_t)w * p.ctx_stride);                            \
        double* coef_sm = reinterpret_cast<double*>(vlr_smem + (size_t)WARPS_PER_CTA * p.ctx_stride) +            \
                          (size_t)w * p.sm_reads * 4;                                                             \
        WarpWs* ws = p.ws + gw;                                                                                   \
        double* coef = p.coef + (int64_t)gw * p.coef_cap * 4;                                                     \
        double* be = p.be ? p.be + (int64_t)gw * p.be_stride : nullptr;                                           \
        for (;;) {                                                                                                \
            unsigned long long t = 0;                                                                             \
            if (lane_id() == 0) t = atomicAdd(p.ticket, 1ULL);                                                    \
            t = __shfl_sync(FULL, t, 0, LANES);                                                                   \
            int64_t locus = (int64_t)t;                                                                           \
            if (p.locus_list) {                                                                                   \
                if (t >= (unsigned long long)*p.locus_list_n) break;                                              \
                locus = p.locus_list[t];                                                                          \
            } else if ((int64_t)t >= p.b.n_loci) break;                                                           \
            NS::process_locus(&p.sc, &p.b, &p.r, ws, coef, coef_sm, p.sm_reads, be, p.coef_cap, locus, c);        \
            warp_sync();                                                                                          \
        }                                                                                                         \
    }
VLR_DEFINE_KERNEL(vlr_small)
VLR_DEFINE_KERNEL(vlr_full)

// ---- wavefront pipeline kernels (engine_wave.cuh) ------------------------------------------------------------------
struct WaveParams {
    DevScenario sc;
    DevBatch b;
    DevResults r;
    WavePlan wp;
    vlr_small::WaveBufs wb;
    WarpWs* ws;       // per warp of the prep/finish grid (AFD scratch)
    int64_t sub_lo;   // first locus of the sub-chunk
    int n_sub;        // loci in the sub-chunk
    int want_be;      // AFD requested: log base events
    int debug;        // VLR_WAVE_DEBUG=1: CTA 0 prints per-phase cycle counts of its first group
};
constexpr int WAVE_ROUND_THREADS = vlr_small::W_GROUP * vlr_small::W_MAXT; // 128
constexpr size_t WAVE_ROUND_SMEM = (size_t)vlr_small::W_GROUP * vlr_small::W_SLOT_STRIDE * sizeof(double);

#ifndef VLR_PREP_MIN_CTAS
#define VLR_PREP_MIN_CTAS 3
#endif
// prep, split in three so that each kernel's text fits the instruction caches (engine_wave.cuh "prep")
__global__ void __launch_bounds__(THREADS, VLR_PREP_MIN_CTAS) vlr_wave_pre_kernel(const __grid_constant__ WaveParams p) {
    using namespace vlr_small;
    // (lean context: the pre-pass, the coefficients, a pileup evaluation and locus_tail never touch the tree walk's state;
    // 2.6 instead of 7.7 KB of shared memory per warp leave the SM's shared memory to the resident kernels next to them)
    Ctx& c = *reinterpret_cast<Ctx*>(vlr_smem + (size_t)group_in_cta() * CTX_LEAN);
    for (;;) {
        unsigned long long t = 0;
        if (lane_id() == 0) t = atomicAdd(&p.wb.cnt->ticket[0], 1ULL);
        t = __shfl_sync(FULL, t, 0, LANES);
        if (t >= (unsigned long long)p.n_sub) break;
        wave_pre_locus(&p.sc, &p.b, p.wp, p.wb, p.sub_lo + (int64_t)t, (int)t, p.want_be != 0, c);
        warp_sync();
    }
}

__global__ void __launch_bounds__(256) vlr_wave_lcinit_kernel(const __grid_constant__ WaveParams p) {
    using namespace vlr_small;
    __shared__ unsigned key_n[R_CLASSES * W_KEYS], key_off[R_CLASSES * W_KEYS];
    for (int i = (int)threadIdx.x; i < R_CLASSES * W_KEYS; i += (int)blockDim.x) key_n[i] = p.wb.cnt->rkey_n[i / W_KEYS][i % W_KEYS];
    __syncthreads();
    if (threadIdx.x < R_CLASSES) wave_key_offsets(key_n + threadIdx.x * W_KEYS, key_off + threadIdx.x * W_KEYS);
    __syncthreads();
    const int n_lc = (int)min(p.wb.cnt->n_lc, (unsigned)p.wb.lc_cap);
    const int lane = (int)(threadIdx.x & 31);
    for (int k0 = (int)(blockIdx.x * blockDim.x + (threadIdx.x & ~31u)); k0 < n_lc; k0 += (int)(gridDim.x * blockDim.x)) {
        const int k = k0 + lane;
        const int id = k < n_lc ? wave_lc_init(&p.sc, p.wp, p.wb, k) : -1;
        // the warp's resident lcs go into their key's block of the class's list: the lcs of one key as one run, in lc
        // order (one atomic per warp and key)
        __syncwarp();
        const unsigned m = __match_any_sync(0xffffffffu, id);
        if (id >= 0) {
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&p.wb.cnt->rkey_cur[id / W_KEYS][id % W_KEYS], (unsigned)__popc(m));
            base = __shfl_sync(m, base, leader);
            const unsigned at = key_off[id] + base + (unsigned)__popc(m & ((1u << lane) - 1u));
            if (at < (unsigned)p.wb.lc_cap) p.wb.rlist[(size_t)(id / W_KEYS) * p.wb.lc_cap + at] = k;
        }
    }
}

__global__ void __launch_bounds__(THREADS, VLR_PREP_MIN_CTAS) vlr_wave_coef_kernel(const __grid_constant__ WaveParams p) {
    using namespace vlr_small;
    Ctx& c = *reinterpret_cast<Ctx*>(vlr_smem + (size_t)group_in_cta() * CTX_LEAN);
#ifdef VLR_COEF_MEMO // experiment: the MAPQ table of vlr_sets_lc_kernel here too - measured neutral (7.199 against 7.194 M loci/s)
    __shared__ MemoTab memo_tabs[WARPS_PER_CTA];
    MemoTab* const memo = &memo_tabs[group_in_cta()];
    memo_clear(memo);
#else
    MemoTab* const memo = nullptr;
#endif
    const unsigned long long n_lc = min(p.wb.cnt->n_lc, (unsigned)p.wb.lc_cap);
    for (;;) {
        unsigned long long t = 0;
        if (lane_id() == 0) t = atomicAdd(&p.wb.cnt->ticket[3], 1ULL);
        t = __shfl_sync(FULL, t, 0, LANES);
        if (t >= n_lc) break; // (phasing the CTA's warps like vlr_sets_lc_kernel: -1 % on config 2, -12 % on config 5's depth skew)
        wave_lc_coef(&p.sc, &p.b, p.wp, p.wb, (int)t, p.sub_lo, p.want_be != 0, c, (int)(blockIdx.x * WARPS_PER_CTA) + group_in_cta(), memo);
        warp_sync();
    }
}

// One CTA per group of lcs of the round's list. A group is processed in passes: pass 0 takes all lcs whose pileups fit
// a coefficient slot (<= W_SLOT_READS reads) together; every deeper lc then gets a pass of its own with the whole slot
// region as one big slot and up to 32 lanes per task, so a 2000-read pileup next to 10-read ones (config 5) neither
// idles the other lanes nor falls back to global-memory coefficient loads. A pass: cp.async the parent sample's
// coefficients into shared memory, every task evaluates its parent pileup; cp.async the leaf sample's coefficients,
// H = 1..32 neighbouring lanes run one task (tasks of an lc sit in neighbouring lanes: their coefficient loads are
// shared-memory broadcasts); then 8-lane groups close the lcs of the pass: trapezoids over the task grids, MAP
// bookkeeping, the next round's tasks.
#ifndef VLR_ROUND_MIN_CTAS
#define VLR_ROUND_MIN_CTAS 4
#endif
constexpr int WAVE_BIG_READS = vlr_small::W_GROUP * vlr_small::W_SLOT_STRIDE / 4; // reads the whole slot region holds
__global__ void __launch_bounds__(WAVE_ROUND_THREADS, VLR_ROUND_MIN_CTAS) vlr_wave_round_kernel(const __grid_constant__ WaveParams p, int round) {
    using namespace vlr_small;
    __shared__ int s_lc[W_GROUP], s_cnt[W_GROUP], s_deep[W_GROUP], s_h[W_GROUP], s_off[W_GROUP + 1], s_ndeep;
    const WaveBufs& wb = p.wb;
    const int n_list = (int)wb.cnt->dlist_n[round]; // the round's lcs with a pileup deeper than a slot
    const int* list = wb.dlist[round & 1];
    WaveTask* tasks = wb.tasks[round & 1];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* slots = reinterpret_cast<double*>(vlr_smem);
    // full groups while the list is long; in the straggler rounds every CTA takes a short group and more lanes per task
    int G = (n_list + (int)gridDim.x - 1) / (int)gridDim.x;
    G = G < 1 ? 1 : (G > W_GROUP ? W_GROUP : G);
    for (int g0 = (int)blockIdx.x * G; g0 < n_list; g0 += (int)gridDim.x * G) {
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
        __syncthreads();
        if (p.debug) tk0 = clock64();
        if (tid < 32) { // warp 0: the group's lcs, their task counts, which of them are deep; thread ranges of pass 0
            const int lci = (tid < G && g0 + tid < n_list) ? list[g0 + tid] : -1;
            int cnt = 0;
            bool deep = false;
            if (lci >= 0) {
                const WaveLC& L = wb.lcs[lci];
                cnt = L.task_count;
                deep = L.nT > W_SLOT_READS || L.nP > W_SLOT_READS;
            }
            const unsigned dm = __ballot_sync(0xffffffffu, deep);
            // lanes per task: a function of the lc alone (its task count; in a deep pass the lc has the CTA to itself),
            // never of the group it happens to share a CTA with, so results are bitwise reproducible whatever order the
            // atomics built the round's list in. Thread ranges are padded to even sizes so that the lane pairs of a task
            // stay aligned for the xor butterflies.
            const int hh = cnt <= 1 ? 8 : (cnt <= 2 ? 4 : (cnt <= 4 ? 2 : 1)); // <= 8 threads per lc
            const int thr = (lci >= 0 && !deep) ? ((cnt * hh + 1) & ~1) : 0;
            int incl = thr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (tid < W_GROUP) {
                s_lc[tid] = lci;
                s_cnt[tid] = cnt;
                s_off[tid] = incl - thr;
                s_h[tid] = hh;
                if (tid == W_GROUP - 1) s_off[W_GROUP] = incl;
                if (deep) s_deep[__popc(dm & ((1u << tid) - 1u))] = tid;
            }
            if (tid == 0) s_ndeep = __popc(dm);
        }
        __syncthreads();
        const int n_deep = s_ndeep;
        for (int pass = 0; pass <= n_deep; ++pass) {
            const int g_deep = pass > 0 ? s_deep[pass - 1] : -1;
            if (pass > 0) { // a deep lc alone: as many lanes per task as fit the CTA
                __syncthreads();
                if (tid < 32) {
                    const bool member = tid == g_deep;
                    int hh = 1;
                    if (member)
                        while (hh < 32 && s_cnt[tid] * hh * 2 <= WAVE_ROUND_THREADS) hh <<= 1;
                    const int thr = member ? ((s_cnt[tid] * hh + 1) & ~1) : 0;
                    int incl = thr;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (tid < W_GROUP) {
                        s_off[tid] = incl - thr;
                        s_h[tid] = hh;
                        if (tid == W_GROUP - 1) s_off[W_GROUP] = incl;
                    }
                }
                __syncthreads();
            }
            const int total = s_off[W_GROUP]; // threads of the pass
            if (total == 0) continue; // (uniform: every thread reads the same shared value)
            const int slot_reads = pass == 0 ? W_SLOT_READS : WAVE_BIG_READS;
            int g_mine = 0, q = 0; // my lc of the group and my task of that lc
            bool active = tid < total;
            WSplit sp;
            sp.H = 1;
            sp.h = 0;
            sp.mask = 1u << lane;
            if (active) {
                while (s_off[g_mine + 1] <= tid) ++g_mine; // (non-members have empty ranges and are skipped)
                const int H = s_h[g_mine], local = tid - s_off[g_mine];
                q = local / H;
                active = q < s_cnt[g_mine]; // (padding lanes idle)
                sp.H = H;
                sp.h = local & (H - 1);
                sp.mask = H == 32 ? 0xffffffffu : (((1u << H) - 1u) << (lane & ~(H - 1)));
            }
            auto slot_of = [&](int g) { return slots + (pass == 0 ? (size_t)g * W_SLOT_STRIDE : (size_t)0); };
            auto stage = [&](bool parent) {
                for (int g = warp; g < G; g += WAVE_ROUND_THREADS / 32) {
                    if (s_off[g + 1] == s_off[g]) continue; // not a member of this pass (or no tasks)
                    const WaveLC& L = wb.lcs[s_lc[g]];
                    const int nr = parent ? L.nP : L.nT;
                    if (nr > slot_reads) continue; // does not even fit the whole region: read from the arena (L2)
                    const double2* src = reinterpret_cast<const double2*>(wb.coef + (parent ? L.coefP : L.coefT));
                    double2* dst = reinterpret_cast<double2*>(slot_of(g));
                    if (pass == 0) {
                        for (int i = lane; i < nr * 2; i += 32) __pipeline_memcpy_async(dst + i, src + i, sizeof(double2));
                    }
                }
                if (pass > 0) { // one lc, all warps copy
                    const WaveLC& L = wb.lcs[s_lc[g_deep]];
                    const int nr = parent ? L.nP : L.nT;
                    if (nr <= slot_reads) {
                        const double2* src = reinterpret_cast<const double2*>(wb.coef + (parent ? L.coefP : L.coefT));
                        double2* dst = reinterpret_cast<double2*>(slots);
                        for (int i = tid; i < nr * 2; i += WAVE_ROUND_THREADS) __pipeline_memcpy_async(dst + i, src + i, sizeof(double2));
                    }
                }
                __pipeline_commit();
                __pipeline_wait_prior(0);
                __syncthreads();
            };
            // ---- the parent sample's coefficients -> slots; every task evaluates its parent pileup once
            stage(true);
            double lh_const = 0.0;
            if (active) {
                const WaveLC& L = wb.lcs[s_lc[g_mine]];
                const bool in_sm = L.nP <= slot_reads;
                const double2* coP = in_sm ? reinterpret_cast<const double2*>(slot_of(g_mine))
                                           : reinterpret_cast<const double2*>(wb.coef + L.coefP);
                lh_const = wave_task_parent(L, tasks[L.task_base + q], coP, in_sm, sp);
            }
            __syncthreads();
            // ---- the leaf sample's coefficients -> slots; the adaptive integrations
            stage(false);
            if (p.debug) tk1 = clock64();
            if (active) {
                const WaveLC& L = wb.lcs[s_lc[g_mine]];
                WaveTask& t = tasks[L.task_base + q];
                const bool in_sm = L.nT <= slot_reads;
                const double2* coT = in_sm ? reinterpret_cast<const double2*>(slot_of(g_mine))
                                           : reinterpret_cast<const double2*>(wb.coef + L.coefT);
                const size_t row = (size_t)(blockIdx.x * blockDim.x) + (size_t)(s_off[g_mine] + q);
                wave_task_run(&p.sc, p.wp, L, t, coT, in_sm, lh_const, wb.gx + row * W_GCAP, wb.gf + row * W_GCAP, sp);
            }
            __syncthreads();
            if (p.debug) tk2 = clock64();
            // the coefficient slots are dead now: every 8-lane group takes 3 x W_GCAP doubles of them as sort scratch
            // and closes one lc of the pass (their global-memory latencies overlap)
            {
                const int gi = tid >> 3;
                WGroup grp;
                grp.lane = tid & 7;
                grp.n = 8;
                grp.mask = 0xffu << (lane & ~7);
                double* scratch = slots + (size_t)gi * 3 * W_GCAP;
                if (gi < G && s_off[gi + 1] > s_off[gi]) {
                    const size_t row0 = (size_t)(blockIdx.x * blockDim.x) + (size_t)s_off[gi];
                    wave_lc_advance(p.wp, wb, s_lc[gi], round, wb.gx + row0 * W_GCAP, wb.gf + row0 * W_GCAP, W_GCAP, scratch,
                                    p.want_be != 0, grp);
                }
            }
            if (p.debug) {
                __syncthreads();
                tk3 = clock64();
                if (blockIdx.x == 0 && tid == 0 && g0 == 0)
                    printf("round %d pass %d: n_list %d, G %d, threads %d: stage+parent %lld, tasks %lld, advance %lld cycles\n",
                           round, pass, n_list, G, total, tk1 - tk0, tk2 - tk1, tk3 - tk2);
            }
        }
    }
}

// The round kernel of the common case (every pileup of the lc fits a slot): no CTA-wide phases at all. A warp owns
// four coefficient slots and takes four lcs of the round's list at a time; the 8 lanes of an octet own one lc through
// all of its phases — cp.async its parent coefficients into the octet's slot, the lc's tasks (H = 8 / #tasks lanes each)
// evaluate their parent pileup, cp.async the leaf coefficients, run the adaptive integrations, then close the lc
// (trapezoids in the slot, MAP, next round's tasks) as the same 8-lane group. Only __syncwarp between the phases: the
// 16 warps of an SM drift apart and overlap each other's copy, compute and bookkeeping phases (the CTA-phase kernel
// lost ~1/3 of its issue slots to barrier waits).
__global__ void __launch_bounds__(WAVE_ROUND_THREADS, VLR_ROUND_MIN_CTAS) vlr_wave_round_warp_kernel(const __grid_constant__ WaveParams p, int round) {
    using namespace vlr_small;
    const WaveBufs& wb = p.wb;
    const int n_list = (int)wb.cnt->list_n[round];
    const int* list = wb.list[round & 1];
    WaveTask* tasks = wb.tasks[round & 1];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int oct = lane >> 3, l8 = lane & 7;
    constexpr int WARPS = WAVE_ROUND_THREADS / 32;
    static_assert(W_GROUP == 4 * WARPS, "one slot per octet");
    double* slot = reinterpret_cast<double*>(vlr_smem) + (size_t)(4 * warp + oct) * W_SLOT_STRIDE;
    const size_t row0 = (size_t)(blockIdx.x * blockDim.x) + (size_t)(warp * 32 + oct * 8); // task rows of my octet
    WGroup grp;
    grp.lane = l8;
    grp.n = 8;
    grp.mask = 0xffu << (oct * 8);
    const int warp_g = (int)blockIdx.x * WARPS + warp, n_warps = (int)gridDim.x * WARPS;
    for (int g0 = warp_g * 4; g0 < n_list; g0 += n_warps * 4) {
        const int lci = g0 + oct < n_list ? list[g0 + oct] : -1;
        int cnt = 0, nP = 0, nT = 0, task_base = 0;
        const double2 *srcP = nullptr, *srcT = nullptr;
        if (lci >= 0) {
            const WaveLC& L = wb.lcs[lci];
            cnt = L.task_count;
            nP = L.nP;
            nT = L.nT;
            task_base = L.task_base;
            srcP = reinterpret_cast<const double2*>(wb.coef + L.coefP);
            srcT = reinterpret_cast<const double2*>(wb.coef + L.coefT);
        }
        // lanes per task: a function of the lc alone (bitwise reproducible results), <= 8 threads per lc
        const int H = cnt <= 1 ? 8 : (cnt <= 2 ? 4 : (cnt <= 4 ? 2 : 1));
        const int q = l8 / H;
        const bool active = lci >= 0 && q < cnt;
        WSplit sp;
        sp.H = H;
        sp.h = l8 & (H - 1);
        sp.mask = ((1u << H) - 1u) << (lane & ~(H - 1));
        __syncwarp(); // the previous lcs' scratch use of the slots is over
        for (int i = l8; i < nP * 2; i += 8) __pipeline_memcpy_async(reinterpret_cast<double2*>(slot) + i, srcP + i, sizeof(double2));
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncwarp();
        double lh_const = 0.0;
        if (active) lh_const = wave_task_parent(wb.lcs[lci], tasks[task_base + q], reinterpret_cast<const double2*>(slot), true, sp);
        __syncwarp();
        for (int i = l8; i < nT * 2; i += 8) __pipeline_memcpy_async(reinterpret_cast<double2*>(slot) + i, srcT + i, sizeof(double2));
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncwarp();
        if (active)
            wave_task_run(&p.sc, p.wp, wb.lcs[lci], tasks[task_base + q], reinterpret_cast<const double2*>(slot), true, lh_const,
                          wb.gx + (row0 + q) * W_GCAP, wb.gf + (row0 + q) * W_GCAP, sp);
        __syncwarp();
        if (lci >= 0)
            wave_lc_advance(p.wp, wb, lci, round, wb.gx + row0 * W_GCAP, wb.gf + row0 * W_GCAP, W_GCAP, slot, p.want_be != 0, grp);
    }
}

// ---- the lc-resident round kernels (engine_resident.cuh) ------------------------------------------------------------
// Persistent grid; a group of G lanes takes an lc by ticket, bulk-copies its pileup polynomials into its shared-memory
// slot (cp.async.bulk, completion on the slot's mbarrier) and runs every round of the lc: (1) the lc's tasks,
// H = G / #tasks lanes each — parent pileup, then the adaptive search of the leaf allele frequency to completion;
// (2) the group closes the round: trapezoids, MAP, the outer integration's next abscissae = the next round's tasks.
// Only __syncwarp between the phases; a group whose lc is complete takes the next one while its neighbours go on.
// G = 8 (an octet per lc, four lcs per warp) serves lcs of size class 1 (both pileups in 48 polynomials = 240 reads);
// G = 32 (a warp per lc) the deeper classes 2..4 (up to ~1000, ~2000 and ~4100 reads: config 5's depth skew), where the
// evaluation itself (reads / H polynomial blocks per lane) outweighs the bookkeeping.
constexpr int RES_THREADS = 64; // two warps: eight octets (36 KB of shared memory, six CTAs per SM) or two warp groups
#ifndef VLR_RES_MIN_CTAS
#define VLR_RES_MIN_CTAS 6
#endif
constexpr size_t res_smem(int G, int slot_q) {
    return (size_t)(RES_THREADS / G) * (sizeof(vlr_small::ROct) + (size_t)slot_q * vlr_small::R_QW * sizeof(double));
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int G>
__device__ __forceinline__ void wave_resident_body(const WaveParams& p, const int cls, const int slot_q) {
    using namespace vlr_small;
    const WaveBufs& wb = p.wb;
    constexpr int GROUPS = RES_THREADS / G; // per CTA
    const size_t gstride = sizeof(ROct) + (size_t)slot_q * R_QW * sizeof(double);
    const int tid = (int)threadIdx.x, lane = tid & 31, gi = tid / G, lg = tid % G;
    ROct& oc = *reinterpret_cast<ROct*>(vlr_smem + (size_t)gi * gstride);
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    WGroup grp;
    grp.lane = lg;
    grp.n = G;
    grp.mask = gmask;
    const size_t og = (size_t)blockIdx.x * GROUPS + (size_t)gi; // this group's rows of the global scratch
    double* const rows_x = wb.rgx + og * (W_MAXT * W_GCAP);
    double* const rows_m = wb.rgm + og * (W_MAXT * W_GCAP);
    int* const rows_e = wb.rge + og * (W_MAXT * W_GCAP);
    double* const big = wb.rscratch + og * (3 * W_GCAP);
    const int* rlist = wb.rlist + (size_t)(cls - 1) * wb.lc_cap;
    const unsigned n_list = min(wb.cnt->rlist_total[cls - 1], (unsigned)wb.lc_cap);
    const unsigned bar = smem_u32(&oc.bar);
    if (lg == 0) {
        oc.q = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(&oc) + sizeof(ROct));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned parity = 0;
    bool have = false, more = true; // more: the list may still hold an lc for this group
    int round = 0, cnt = 0;
    for (;;) {
        if (!have && more) {
            unsigned long long t = 0;
            if (lg == 0) t = atomicAdd(&wb.cnt->ticket[4 + cls - 1], 1ULL);
            t = __shfl_sync(gmask, t, 0, G);
            more = t < (unsigned long long)n_list;
            const int lci = more ? rlist[t] : -1;
            if (lci >= 0) { // (a negative entry: the list's bookkeeping left a hole; take the next ticket)
                const WaveLC& L = wb.lcs[lci];
                const int nqP = L.nqPx + L.nqPy, nq = L.nqTx + L.nqTy; // the slot takes the leaf pileup's polynomials
                __syncwarp(gmask); // nobody of the group still reads the slot
                if (lg == 0) {
                    oc.lc.lci = lci;
                    oc.lc.li = L.li;
                    oc.lc.ci = L.ci;
                    oc.lc.nqPx = L.nqPx;
                    oc.lc.nqPy = L.nqPy;
                    oc.lc.nqTx = L.nqTx;
                    oc.lc.nqTy = L.nqTy;
                    oc.lc.ksumP = L.ksumP;
                    oc.lc.ksumT = L.ksumT;
                    oc.lc.qP = wb.coef + L.coefP;
                    const unsigned bytes = (unsigned)nq * (unsigned)(R_QW * sizeof(double));
                    // the slot was last read through the generic proxy: order those reads before the bulk write
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(oc.q)),
                                 "l"(wb.coef + L.coefP + (size_t)nqP * R_QW), "r"(bytes), "r"(bar)
                                 : "memory");
                }
                cnt = r_first_tasks(&p.sc, p.wp, wb, lci, oc.task, grp); // while the copy is in flight
                unsigned ok = 0;
                while (!ok) {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                                 : "=r"(ok)
                                 : "r"(bar), "r"(parity)
                                 : "memory");
                }
                parity ^= 1u;
                have = true;
                round = 0;
            }
        }
#ifdef VLR_RES_CTA_SYNC // experiment: the CTA's warps enter the task and advance phases together (instruction caches)
        if (!__syncthreads_or((have || more) ? 1 : 0)) break;
#else
        if (!__any_sync(0xffffffffu, have || more)) break;
        __syncwarp();
#endif
        {
            // lanes per task: a function of the lc's class and task count alone (bitwise reproducible results)
            const int H = cnt <= 1 ? G : (cnt <= 2 ? G / 2 : (cnt <= 4 ? G / 4 : G / 8));
            const int q = lg / H;
            const bool runs = have && q < cnt;
            const unsigned wmask = __ballot_sync(0xffffffffu, runs); // the warp's task lanes of this round
            if (runs) {
                WSplit sp;
                sp.H = H;
                sp.h = lg & (H - 1);
                sp.mask = H == 32 ? 0xffffffffu : (((1u << H) - 1u) << (lane & ~(H - 1)));
                r_task(&p.sc, p.wp, wb, oc, q, cnt, rows_x + (size_t)q * W_GCAP, rows_m + (size_t)q * W_GCAP, rows_e + (size_t)q * W_GCAP, sp, wmask);
            }
        }
#ifdef VLR_RES_CTA_SYNC
        __syncthreads();
#else
        __syncwarp();
#endif
        if (have) {
            const int next = r_advance(&p.sc, p.wp, wb, oc, round, cnt, rows_x, rows_m, rows_e, big, p.want_be != 0, grp);
            if (next == 0) have = false;
            cnt = next;
            round++;
        }
    }
}

__global__ void __launch_bounds__(RES_THREADS, VLR_RES_MIN_CTAS) vlr_wave_resident_kernel(const __grid_constant__ WaveParams p) {
    wave_resident_body<8>(p, 1, vlr_small::R_SLOT_Q);
}
// size classes 2..4 (`cls`): a warp per lc, `slot_q` polynomials per slot
__global__ void __launch_bounds__(RES_THREADS, 2) vlr_wave_resident_deep_kernel(const __grid_constant__ WaveParams p, int cls, int slot_q) {
    wave_resident_body<32>(p, cls, slot_q);
}

#ifndef VLR_FIN_MIN_CTAS
#define VLR_FIN_MIN_CTAS 3 // (80 registers, no spills; 2 -> 3 CTAs per SM: +0.6 % on config 2)
#endif
__global__ void __launch_bounds__(THREADS, VLR_FIN_MIN_CTAS) vlr_wave_finish_kernel(const __grid_constant__ WaveParams p) {
    using namespace vlr_small;
    Ctx& c = *reinterpret_cast<Ctx*>(vlr_smem + (size_t)group_in_cta() * CTX_LEAN);
    WarpWs* ws = p.ws + (blockIdx.x * WARPS_PER_CTA + group_in_cta());
    for (;;) {
        unsigned long long t = 0;
        if (lane_id() == 0) t = atomicAdd(&p.wb.cnt->ticket[1], 1ULL);
        t = __shfl_sync(FULL, t, 0, LANES);
        if (t >= (unsigned long long)p.n_sub) break;
        wave_finish_locus(&p.sc, &p.b, &p.r, p.wp, p.wb, ws, p.sub_lo + (int64_t)t, (int)t, c);
        warp_sync();
    }
}

// ---- all-Set pipeline kernels (engine_sets.cuh) -----------------------------------------------------------------------
struct SetsParams {
    DevScenario sc;
    DevBatch b;
    DevResults r;
    SetsPlan sp;
    vlr_small::SetsBufs sb;
    WarpWs* ws;
    int64_t sub_lo;
    int n_sub;
    int want_be;
};
constexpr size_t SETS_LC_WARP_SMEM = (size_t)vlr_small::CTX_LEAN + sizeof(double) * (4 * vlr_small::SETS_SM_READS + SETS_MAXF);

__global__ void __launch_bounds__(THREADS, VLR_PREP_MIN_CTAS) vlr_sets_pre_kernel(const __grid_constant__ SetsParams p) {
    using namespace vlr_small;
    Ctx& c = *reinterpret_cast<Ctx*>(vlr_smem + (size_t)group_in_cta() * CTX_LEAN);
    for (;;) {
        unsigned long long t = 0;
        if (lane_id() == 0) t = atomicAdd(&p.sb.cnt->ticket[0], 1ULL);
        t = __shfl_sync(FULL, t, 0, LANES);
        if (t >= (unsigned long long)p.n_sub) break;
        sets_pre_locus(&p.sc, &p.b, p.sb, p.sub_lo + (int64_t)t, (int)t, p.want_be != 0, c);
        warp_sync();
    }
}

// warp per lc; shared memory per warp: [Ctx][coefficient arena: SETS_SM_READS x 4 doubles][fold values: SETS_MAXF doubles]
__global__ void __launch_bounds__(THREADS, 2) vlr_sets_lc_kernel(const __grid_constant__ SetsParams p) {
    using namespace vlr_small;
    Ctx& c = *reinterpret_cast<Ctx*>(vlr_smem + (size_t)group_in_cta() * CTX_LEAN);
    double* arena = reinterpret_cast<double*>(vlr_smem + (size_t)WARPS_PER_CTA * CTX_LEAN) +
                    (size_t)group_in_cta() * (4 * SETS_SM_READS + SETS_MAXF);
    __shared__ MemoTab memo_tabs[WARPS_PER_CTA]; // ln(1 - e^{prob_mapping}) per MAPQ value, one table per warp (read_coefficients)
    MemoTab* const memo = &memo_tabs[group_in_cta()];
    memo_clear(memo);
    const unsigned long long n_lc = min(p.sb.cnt->n_lc, (unsigned)p.sb.lc_cap);
    for (;;) { // the CTA's warps take one lc each and go through its phases together (sets_lc)
        unsigned long long t = 0;
        if (lane_id() == 0) t = atomicAdd(&p.sb.cnt->ticket[1], 1ULL);
        t = __shfl_sync(FULL, t, 0, LANES);
        const bool mine = t < n_lc;
        if (!__syncthreads_or(mine ? 1 : 0)) break;
        sets_lc(&p.sc, &p.b, p.sp, p.sb, mine ? (int)t : -1, p.sub_lo, p.want_be != 0, c, arena, arena + 4 * SETS_SM_READS, true, memo);
        warp_sync();
    }
}

__global__ void __launch_bounds__(THREADS, 2) vlr_sets_finish_kernel(const __grid_constant__ SetsParams p) {
    using namespace vlr_small;
    Ctx& c = *reinterpret_cast<Ctx*>(vlr_smem + (size_t)group_in_cta() * CTX_LEAN);
    WarpWs* ws = p.ws + (blockIdx.x * WARPS_PER_CTA + group_in_cta());
    for (;;) {
        unsigned long long t = 0;
        if (lane_id() == 0) t = atomicAdd(&p.sb.cnt->ticket[2], 1ULL);
        t = __shfl_sync(FULL, t, 0, LANES);
        if (t >= (unsigned long long)p.n_sub) break;
        sets_finish_locus(&p.sc, &p.b, &p.r, p.sp, p.sb, ws, p.sub_lo + (int64_t)t, (int)t, c);
        warp_sync();
    }
}

// fp64 peak microbenchmark: 8 independent FMA chains per thread, register resident
__global__ void __launch_bounds__(256) vlr_fp64_peak_kernel(double* out, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3,
           a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

inline int align16(size_t x) { return (int)((x + 15) & ~(size_t)15); }

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) return ctx->fail_cuda(_e, #call, __LINE__);            \
    } while (0)

// ---- packed batch (include/vlr_engine.h: vlr_packed_batch_t) --------------------------------------------------------
// Widens the encoded columns of one chunk to the f32 / u32 columns the pipelines read: blockIdx.y = column, four rows
// per thread (one 4- or 8-byte load, one 16-byte store, both coalesced). HBM-bound: <= 9 bytes in, 32 bytes out per read.
struct UnpackParams {
    const void* src[VLR_N_PACKED_COLUMNS];
    const uint32_t* dict[VLR_N_PACKED_COLUMNS];
    uint32_t* dst[VLR_N_PACKED_COLUMNS];
    int enc[VLR_N_PACKED_COLUMNS];
    int n_dict[VLR_N_PACKED_COLUMNS];
    int64_t n;
};

__global__ void __launch_bounds__(256) vlr_unpack_kernel(const __grid_constant__ UnpackParams p) {
    __shared__ uint32_t sdict[256];
    const int c = (int)blockIdx.y;
    const int enc = p.enc[c];
    if (enc == VLR_ENC_F32) return; // copied straight into place
    uint32_t* __restrict__ dst = p.dst[c];
    const int64_t n = p.n, n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (enc == VLR_ENC_DICT8) {
        for (int i = (int)threadIdx.x; i < 256; i += (int)blockDim.x) sdict[i] = i < p.n_dict[c] ? p.dict[c][i] : 0u;
        __syncthreads();
        const uchar4* __restrict__ s4 = reinterpret_cast<const uchar4*>(p.src[c]);
        for (int64_t i = t0; i < n4; i += stride) {
            const uchar4 k = s4[i];
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(sdict[k.x], sdict[k.y], sdict[k.z], sdict[k.w]);
        }
        const uint8_t* __restrict__ s1 = reinterpret_cast<const uint8_t*>(p.src[c]);
        for (int64_t i = (n4 << 2) + t0; i < n; i += stride) dst[i] = sdict[s1[i]];
    } else if (enc == VLR_ENC_DICT16) {
        const uint32_t* __restrict__ d = p.dict[c];
        const ushort4* __restrict__ s4 = reinterpret_cast<const ushort4*>(p.src[c]);
        for (int64_t i = t0; i < n4; i += stride) {
            const ushort4 k = s4[i];
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(__ldg(d + k.x), __ldg(d + k.y), __ldg(d + k.z), __ldg(d + k.w));
        }
        const uint16_t* __restrict__ s1 = reinterpret_cast<const uint16_t*>(p.src[c]);
        for (int64_t i = (n4 << 2) + t0; i < n; i += stride) dst[i] = __ldg(d + s1[i]);
    } else if (enc == VLR_ENC_F16) {
        const ushort4* __restrict__ s4 = reinterpret_cast<const ushort4*>(p.src[c]);
        auto widen = [](unsigned short h) { return __float_as_uint(__half2float(__ushort_as_half(h))); };
        for (int64_t i = t0; i < n4; i += stride) {
            const ushort4 k = s4[i];
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(widen(k.x), widen(k.y), widen(k.z), widen(k.w));
        }
        const uint16_t* __restrict__ s1 = reinterpret_cast<const uint16_t*>(p.src[c]);
        for (int64_t i = (n4 << 2) + t0; i < n; i += stride) dst[i] = widen(s1[i]);
    } else { // VLR_ENC_CONST
        const uint32_t v = p.dict[c][0];
        for (int64_t i = t0; i < n4; i += stride) reinterpret_cast<uint4*>(dst)[i] = make_uint4(v, v, v, v);
        for (int64_t i = (n4 << 2) + t0; i < n; i += stride) dst[i] = v;
    }
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Slot { // one in-flight chunk of vlr_call_batch
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_free = nullptr; // inputs of the chunk are on the device / its kernels have read them
    DevBuf offsets, cols[7], rflags, hart, hvar, lflags, het, semr;
    DevBuf pk[VLR_N_PACKED_COLUMNS], pk_dict[VLR_N_PACKED_COLUMNS]; // packed chunk (vlr_call_batch_packed) and its dictionaries
    DevBuf log_post, log_marginal, map_vaf, map_config, best_event, status, n_base, afd_count, afd_vaf, afd_logp;
    DevBuf ws, coef, be, ticket;
    int coef_cap = 0;
    bool be_ready = false;
    // wavefront pipeline workspace (one sub-chunk of loci at a time)
    DevBuf w_cnt, w_loci, w_lcs, w_ogx, w_ogf, w_coef, w_tasks[2], w_list[2], w_dlist[2], w_deferred, w_gx, w_gf, w_be, w_ben;
    DevBuf w_rlist, w_rgx, w_rgm, w_rge, w_rscratch, w_cscratch; // lc-resident round kernel
    DevBuf s_cnt, s_loci, s_lcs, s_deferred, s_be, s_ben;        // all-Set pipeline
};

} // namespace

struct vlr_ctx {
    int device = 0;
    int n_sms = 0;
    int ctas_per_sm = 1;
    int grid = 0;
    int S = 0, E = 0;
    bool small = false;   // which engine variant serves this scenario
    bool wave = false;    // two-level chain scenario: the wavefront pipeline serves it (deferring loci it cannot)
    WavePlan wplan;
    int wave_grid_prep = 0, wave_grid_round = 0, wave_grid_finish = 0, wave_grid_res = 0;
    int wave_grid_deep[3] = {0, 0, 0}; // resident kernels of size classes 2..4 (a warp per lc)
    bool resident = true; // VLR_RESIDENT=0: per-round kernels only (A/B measurements)
    bool sets = false;    // all-Set scenario (pedigrees): the all-Set pipeline serves it (engine_sets.cuh)
    SetsPlan splan;
    int sets_grid_pre = 0, sets_grid_lc = 0, sets_grid_finish = 0;
    size_t sets_smem_lc = 0;
    DevBuf d_sets_folds, d_sets_leaves, d_sets_leaf_vaf, d_sets_prior_val, d_sets_prior_side, d_sets_prior_state;
    size_t wave_smem_prep = 0;
    int ctx_stride = 0;
    size_t smem_bytes = 0;
    ScenarioPrep prep;
    DevScenario dsc;
    DevBuf d_samples, d_events, d_nodes, d_set_vafs, d_spectra, d_lfc_nodes, d_lfc_ordinal, d_prior_tab;
    cudaStream_t stream = nullptr; // the context's own stream (device-pointer entry)
    Slot dev_slot;                 // workspace of the device-pointer entry
    static constexpr int MAX_AUX = 4;
    Slot dev_slot_aux[MAX_AUX - 1]; // parts of a large wavefront batch run concurrently on internal streams
    cudaStream_t aux[MAX_AUX] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_AUX] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_done = nullptr; // end of the previous device-entry call: calls share one workspace, so a call on
    bool have_done = false;        // another stream waits for it (an event chain instead of a rule for the caller)
    int n_aux = 3; // measured on 512k config-2 loci: 1 stream 4.05, 2: 4.43, 3: 4.48, 4: 4.51 M loci/s
    Slot slots[NBUF_MAX];
    cudaStream_t copy_stream = nullptr; // all host -> device copies of the host entries (prefetch ahead of the kernels)
    int nbuf = 6; // chunk slots in use (VLR_NBUF): nbuf / 2 compute streams + one copy stream that runs ahead of them.
                  // Measured (1M config-2 / config-3 loci, packed columns): 4, 6, 8 slots: 6.83 / 11.9 M loci/s each; before
                  // the copy stream (every slot its own stream for copies and kernels) 3: 6.10 / 9.0, 8: 6.53 / 11.3 — the
                  // streams ran in phase and the SMs idled during every wave of copies
    int64_t reserve_reads = 4096;
    int64_t launches = 0;
    std::string err;

    vlr_status_t fail(vlr_status_t st, const std::string& msg) {
        err = msg;
        return st;
    }
    vlr_status_t fail_cuda(cudaError_t e, const char* what, int line) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at vlr_engine.cu:%d: %s", (int)e, cudaGetErrorString(e), line, what);
        err = buf;
        return e == cudaErrorMemoryAllocation ? VLR_ERR_OUT_OF_MEMORY : VLR_ERR_CUDA;
    }
};

namespace {

vlr_status_t ensure_workspace(vlr_ctx* ctx, Slot& sl, int64_t max_reads, bool want_be) {
    const int64_t n_warps = (int64_t)ctx->grid * WARPS_PER_CTA;
    if (max_reads < 1) max_reads = 1;
    if (max_reads > 0x7fffffff / 8) return ctx->fail(VLR_ERR_UNSUPPORTED, "locus with more than 2^28 reads");
    CK(sl.ws.ensure(sizeof(WarpWs) * (size_t)n_warps));
    if (max_reads > sl.coef_cap) {
        int64_t cap = max_reads + max_reads / 8;
        CK(sl.coef.ensure((size_t)n_warps * (size_t)cap * 4 * sizeof(double)));
        sl.coef_cap = (int)cap;
    }
    if (want_be) CK(sl.be.ensure((size_t)n_warps * BE_CAP * (2 + ctx->S) * sizeof(double)));
    CK(sl.ticket.ensure(sizeof(unsigned long long)));
    return VLR_OK;
}

// Wavefront pipeline over the batch, one sub-chunk of loci after the other on `stream`: prep -> rounds -> finish ->
// generic engine for the deferred loci. `avg_reads` (reads per locus of the batch, rounded up) sizes the coefficient
// arena; loci that do not fit are deferred, never dropped.
vlr_status_t launch_wave(vlr_ctx* ctx, Slot& sl, const DevBatch& b, const DevResults& r, int64_t avg_reads, cudaStream_t stream,
                         int64_t locus_begin, int64_t locus_end) {
    using namespace vlr_small;
    const bool want_be = r.afd_capacity > 0;
    // sub-chunk: <= 131072 loci (8192 with an AFD: the base-event log is 128 KB per locus) and <= ~32M reads, so that
    // deep batches keep the workspace bounded, and no more than the range holds; arena and lc table sized for 6
    // artifact configs per locus on average (config 2 has 2.2, the depth-skewed config 5 ~4); what does not fit is
    // deferred to the generic engine, never lost. Measured on 1 M config-2 loci (three streams): sub-chunks of 65 536
    // loci 7.04, 131 072: 7.30, 262 144 (four streams, 94 GB of workspace): 7.44 M loci/s - every sub-chunk ends in a
    // tail of straggling lcs; config 5 (100 k loci): 5 592 loci per sub-chunk 1.35, 11 184: 1.51, 22 368: 1.51.
    if (avg_reads < 16) avg_reads = 16;
    int n_sub_cap = want_be ? 8192 : 131072;
    n_sub_cap = (int)std::min<int64_t>(n_sub_cap, std::max<int64_t>(4096, ((int64_t)1 << 25) / avg_reads));
    n_sub_cap = (int)std::min<int64_t>(n_sub_cap, std::max<int64_t>(4096, locus_end - locus_begin));
    if (const char* e = getenv("VLR_WAVE_SUB")) { // tuning knob: loci per sub-chunk
        const int v = atoi(e);
        if (v >= 256 && v <= (1 << 20)) n_sub_cap = want_be ? std::min(v, 8192) : v;
    }
    const int lc_cap = n_sub_cap * 6;
    const int64_t coef_cap = ((int64_t)n_sub_cap * avg_reads * 6 + (1 << 20)) * 4; // doubles
    const int g_stride = ctx->wave_grid_round * WAVE_ROUND_THREADS;
    CK(sl.w_cnt.ensure(sizeof(WaveCounters)));
    CK(sl.w_loci.ensure(sizeof(WaveLocus) * (size_t)n_sub_cap));
    CK(sl.w_lcs.ensure(sizeof(WaveLC) * (size_t)lc_cap));
    CK(sl.w_ogx.ensure(sizeof(double) * (size_t)lc_cap * W_OGRID));
    CK(sl.w_ogf.ensure(sizeof(double) * (size_t)lc_cap * W_OGRID));
    CK(sl.w_coef.ensure(sizeof(double) * (size_t)coef_cap));
    {
        // (the deep kernels have two groups per CTA and fewer CTAs: they share the octet kernel's rows)
        const size_t n_oct = (size_t)std::max(ctx->wave_grid_res * (RES_THREADS / 8),
                                              std::max(ctx->wave_grid_deep[0], std::max(ctx->wave_grid_deep[1], ctx->wave_grid_deep[2])) * (RES_THREADS / 32));
        CK(sl.w_rlist.ensure(sizeof(int) * (size_t)lc_cap * R_CLASSES));
        CK(sl.w_rgx.ensure(sizeof(double) * n_oct * W_MAXT * W_GCAP));
        CK(sl.w_rgm.ensure(sizeof(double) * n_oct * W_MAXT * W_GCAP));
        CK(sl.w_rge.ensure(sizeof(int) * n_oct * W_MAXT * W_GCAP));
        CK(sl.w_rscratch.ensure(sizeof(double) * n_oct * 3 * W_GCAP));
    }
    // scratch of the coefficient kernel: as deep as the deepest locus the workspace was sized for (ensure_workspace)
    const int cs_reads = std::min<int>(R_MAXREADS, std::max(sl.coef_cap, 256));
    CK(sl.w_cscratch.ensure(sizeof(double) * (size_t)ctx->wave_grid_prep * WARPS_PER_CTA * 6 * (size_t)cs_reads));
    for (int i = 0; i < 2; ++i) {
        CK(sl.w_tasks[i].ensure(sizeof(WaveTask) * (size_t)lc_cap * W_MAXT));
        CK(sl.w_list[i].ensure(sizeof(int) * (size_t)lc_cap));
        CK(sl.w_dlist[i].ensure(sizeof(int) * (size_t)lc_cap));
    }
    CK(sl.w_deferred.ensure(sizeof(int) * (size_t)n_sub_cap));
    CK(sl.w_gx.ensure(sizeof(double) * (size_t)W_GCAP * g_stride));
    CK(sl.w_gf.ensure(sizeof(double) * (size_t)W_GCAP * g_stride));
    CK(sl.w_ben.ensure(sizeof(unsigned) * (size_t)n_sub_cap));
    if (want_be) CK(sl.w_be.ensure(sizeof(double) * 4 * (size_t)BE_CAP * n_sub_cap));
    WaveParams p;
    p.sc = ctx->dsc;
    p.b = b;
    p.r = r;
    p.wp = ctx->wplan;
    p.wb.cnt = (WaveCounters*)sl.w_cnt.p;
    p.wb.loci = (WaveLocus*)sl.w_loci.p;
    p.wb.lcs = (WaveLC*)sl.w_lcs.p;
    p.wb.og_x = (double*)sl.w_ogx.p;
    p.wb.og_f = (double*)sl.w_ogf.p;
    p.wb.coef = (double*)sl.w_coef.p;
    p.wb.tasks[0] = (WaveTask*)sl.w_tasks[0].p;
    p.wb.tasks[1] = (WaveTask*)sl.w_tasks[1].p;
    p.wb.list[0] = (int*)sl.w_list[0].p;
    p.wb.list[1] = (int*)sl.w_list[1].p;
    p.wb.dlist[0] = (int*)sl.w_dlist[0].p;
    p.wb.dlist[1] = (int*)sl.w_dlist[1].p;
    p.wb.deferred = (int*)sl.w_deferred.p;
    p.wb.gx = (double*)sl.w_gx.p;
    p.wb.gf = (double*)sl.w_gf.p;
    p.wb.be = want_be ? (double*)sl.w_be.p : nullptr;
    p.wb.be_n = (unsigned*)sl.w_ben.p;
    p.wb.coef_cap = coef_cap;
    p.wb.lc_cap = lc_cap;
    p.wb.rlist = (int*)sl.w_rlist.p;
    p.wb.rgx = (double*)sl.w_rgx.p;
    p.wb.rgm = (double*)sl.w_rgm.p;
    p.wb.rge = (int*)sl.w_rge.p;
    p.wb.rscratch = (double*)sl.w_rscratch.p;
    p.wb.cscratch = (double*)sl.w_cscratch.p;
    p.wb.cscratch_reads = cs_reads;
    const bool deep_classes = cs_reads > R_DEG * R_SLOT_Q; // loci deeper than an octet slot can occur
    p.wb.allow_resident = ctx->resident ? 1 : 0;
    p.ws = (WarpWs*)sl.ws.p;
    p.want_be = want_be ? 1 : 0;
    {
        const char* dbg = getenv("VLR_WAVE_DEBUG");
        p.debug = dbg && dbg[0] == '1';
    }
    KernelParams gp; // generic engine over the deferred loci
    gp.sc = ctx->dsc;
    gp.b = b;
    gp.r = r;
    gp.ws = (WarpWs*)sl.ws.p;
    gp.coef = (double*)sl.coef.p;
    gp.be = want_be ? (double*)sl.be.p : nullptr;
    gp.ticket = &p.wb.cnt->ticket[2];
    gp.locus_list = p.wb.deferred;
    gp.locus_list_n = &p.wb.cnt->n_deferred;
    gp.coef_cap = sl.coef_cap;
    gp.sm_reads = SM_READS;
    gp.ctx_stride = ctx->ctx_stride;
    gp.be_stride = (int64_t)BE_CAP * (2 + ctx->S);
    for (int64_t lo = locus_begin; lo < locus_end; lo += n_sub_cap) {
        p.sub_lo = lo;
        p.n_sub = (int)std::min<int64_t>(n_sub_cap, locus_end - lo);
        CK(cudaMemsetAsync(sl.w_cnt.p, 0, sizeof(WaveCounters), stream));
        CK(cudaMemsetAsync(sl.w_rlist.p, 0xff, sizeof(int) * (size_t)lc_cap * R_CLASSES, stream)); // (holes read as -1)
        vlr_wave_pre_kernel<<<ctx->wave_grid_prep, THREADS, ctx->wave_smem_prep, stream>>>(p);
        vlr_wave_lcinit_kernel<<<ctx->n_sms * 4, 256, 0, stream>>>(p);
        vlr_wave_coef_kernel<<<ctx->wave_grid_prep, THREADS, ctx->wave_smem_prep, stream>>>(p);
        if (ctx->resident) {
            vlr_wave_resident_kernel<<<ctx->wave_grid_res, RES_THREADS, res_smem(8, R_SLOT_Q), stream>>>(p);
            if (deep_classes) { // (pileups that deep exist in this batch)
                const int slot_q[3] = {R_SLOT_QM, R_SLOT_QD, R_SLOT_QL};
                for (int cls = 2; cls <= R_CLASSES; ++cls)
                    vlr_wave_resident_deep_kernel<<<ctx->wave_grid_deep[cls - 2], RES_THREADS, res_smem(32, slot_q[cls - 2]), stream>>>(p, cls, slot_q[cls - 2]);
            }
        }
        for (int round = 0; round < ctx->wplan.max_rounds; ++round) {
            vlr_wave_round_warp_kernel<<<ctx->wave_grid_round, WAVE_ROUND_THREADS, WAVE_ROUND_SMEM, stream>>>(p, round);
            vlr_wave_round_kernel<<<ctx->wave_grid_round, WAVE_ROUND_THREADS, WAVE_ROUND_SMEM, stream>>>(p, round);
        }
        vlr_wave_finish_kernel<<<ctx->wave_grid_finish, THREADS, ctx->wave_smem_prep, stream>>>(p);
        vlr_call_kernel_vlr_small<<<ctx->grid, THREADS, ctx->smem_bytes, stream>>>(gp);
        CK(cudaGetLastError());
        ctx->launches += 5 + (ctx->resident ? (deep_classes ? R_CLASSES : 1) : 0) + 2 * ctx->wplan.max_rounds;
    }
    return VLR_OK;
}

// All-Set pipeline over the batch, one sub-chunk of loci after the other on `stream`: pre -> lc -> finish -> generic
// engine for the deferred loci.
vlr_status_t launch_sets(vlr_ctx* ctx, Slot& sl, const DevBatch& b, const DevResults& r, cudaStream_t stream) {
    using namespace vlr_small;
    const bool want_be = r.afd_capacity > 0;
    const int n_sub_cap = 65536;
    const int lc_cap = n_sub_cap * 4;
    const int S = ctx->S;
    CK(sl.s_cnt.ensure(sizeof(SetsCounters)));
    CK(sl.s_loci.ensure(sizeof(SetsLocus) * (size_t)n_sub_cap));
    CK(sl.s_lcs.ensure(sizeof(SetsLC) * (size_t)lc_cap));
    CK(sl.s_deferred.ensure(sizeof(int) * (size_t)n_sub_cap));
    CK(sl.s_ben.ensure(sizeof(unsigned) * (size_t)n_sub_cap));
    if (want_be) CK(sl.s_be.ensure(sizeof(double) * (size_t)n_sub_cap * SETS_MAXL * (2 + S)));
    SetsParams p;
    p.sc = ctx->dsc;
    p.b = b;
    p.r = r;
    p.sp = ctx->splan;
    p.sb.cnt = (SetsCounters*)sl.s_cnt.p;
    p.sb.loci = (SetsLocus*)sl.s_loci.p;
    p.sb.lcs = (SetsLC*)sl.s_lcs.p;
    p.sb.deferred = (int*)sl.s_deferred.p;
    p.sb.be = want_be ? (double*)sl.s_be.p : nullptr;
    p.sb.be_n = (unsigned*)sl.s_ben.p;
    p.sb.lc_cap = lc_cap;
    p.ws = (WarpWs*)sl.ws.p;
    p.want_be = want_be ? 1 : 0;
    KernelParams gp; // generic engine over the deferred loci
    gp.sc = ctx->dsc;
    gp.b = b;
    gp.r = r;
    gp.ws = (WarpWs*)sl.ws.p;
    gp.coef = (double*)sl.coef.p;
    gp.be = want_be ? (double*)sl.be.p : nullptr;
    gp.ticket = &p.sb.cnt->ticket[3];
    gp.locus_list = p.sb.deferred;
    gp.locus_list_n = &p.sb.cnt->n_deferred;
    gp.coef_cap = sl.coef_cap;
    gp.sm_reads = SM_READS;
    gp.ctx_stride = ctx->ctx_stride;
    gp.be_stride = (int64_t)BE_CAP * (2 + ctx->S);
    for (int64_t lo = 0; lo < b.n_loci; lo += n_sub_cap) {
        p.sub_lo = lo;
        p.n_sub = (int)std::min<int64_t>(n_sub_cap, b.n_loci - lo);
        CK(cudaMemsetAsync(sl.s_cnt.p, 0, sizeof(SetsCounters), stream));
        vlr_sets_pre_kernel<<<ctx->sets_grid_pre, THREADS, ctx->wave_smem_prep, stream>>>(p);
        vlr_sets_lc_kernel<<<ctx->sets_grid_lc, THREADS, ctx->sets_smem_lc, stream>>>(p);
        vlr_sets_finish_kernel<<<ctx->sets_grid_finish, THREADS, ctx->wave_smem_prep, stream>>>(p);
        vlr_call_kernel_vlr_small<<<ctx->grid, THREADS, ctx->smem_bytes, stream>>>(gp);
        CK(cudaGetLastError());
        ctx->launches += 4;
    }
    return VLR_OK;
}

vlr_status_t launch(vlr_ctx* ctx, Slot& sl, const DevBatch& b, const DevResults& r, cudaStream_t stream, int64_t avg_reads = 0) {
    if (ctx->wave && b.n_loci > 0) return launch_wave(ctx, sl, b, r, avg_reads, stream, 0, b.n_loci);
    if (ctx->sets && b.n_loci > 0) return launch_sets(ctx, sl, b, r, stream);
    KernelParams p;
    p.sc = ctx->dsc;
    p.b = b;
    p.r = r;
    p.ws = (WarpWs*)sl.ws.p;
    p.coef = (double*)sl.coef.p;
    p.be = r.afd_capacity > 0 ? (double*)sl.be.p : nullptr;
    p.ticket = (unsigned long long*)sl.ticket.p;
    p.locus_list = nullptr;
    p.locus_list_n = nullptr;
    p.coef_cap = sl.coef_cap;
    p.sm_reads = SM_READS;
    p.ctx_stride = ctx->ctx_stride;
    p.be_stride = (int64_t)BE_CAP * (2 + ctx->S);
    CK(cudaMemsetAsync(sl.ticket.p, 0, sizeof(unsigned long long), stream));
    if (b.n_loci > 0) {
        if (ctx->small) vlr_call_kernel_vlr_small<<<ctx->grid, THREADS, ctx->smem_bytes, stream>>>(p);
        else vlr_call_kernel_vlr_full<<<ctx->grid, THREADS, ctx->smem_bytes, stream>>>(p);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    return VLR_OK;
}

bool results_valid(const vlr_results_t* r) {
    if (!r || !r->log_posteriors || !r->map_vaf || !r->status) return false;
    if (r->afd_capacity < 0) return false;
    if (r->afd_capacity > 0 && (!r->afd_count || !r->afd_vaf || !r->afd_logp)) return false;
    return true;
}
bool batch_valid(const vlr_batch_t* b) {
    if (!b || b->n_loci < 0 || b->n_reads < 0) return false;
    if (b->n_loci == 0) return true;
    if (!b->read_offsets || !b->locus_flags) return false;
    if (b->n_reads > 0 && (!b->prob_mapping || !b->prob_ref || !b->prob_alt || !b->prob_missed_allele ||
                           !b->prob_sample_alt || !b->prob_double_overlap || !b->prob_hit_base || !b->read_flags))
        return false;
    return true;
}

} // namespace

extern "C" {

int32_t vlr_abi_version(void) { return VLR_ABI_VERSION; }

const char* vlr_status_string(vlr_status_t st) {
    switch (st) {
    case VLR_OK: return "ok";
    case VLR_ERR_INVALID_ARGUMENT: return "invalid argument";
    case VLR_ERR_UNSUPPORTED: return "unsupported";
    case VLR_ERR_CUDA: return "CUDA error";
    case VLR_ERR_OUT_OF_MEMORY: return "out of device memory";
    case VLR_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
    }
}

const char* vlr_last_error(const vlr_ctx_t* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t vlr_last_launch_count(const vlr_ctx_t* ctx) { return ctx ? ctx->launches : 0; }
void* vlr_ctx_stream(const vlr_ctx_t* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

vlr_status_t vlr_measure_fp64_peak(int32_t device, double* tflops) {
    if (!tflops) return VLR_ERR_INVALID_ARGUMENT;
    *tflops = 0.0;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return VLR_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n_dev || cudaSetDevice(device) != cudaSuccess) return VLR_ERR_INVALID_ARGUMENT;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VLR_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads) != cudaSuccess) return VLR_ERR_OUT_OF_MEMORY;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) { // first repetition warms up
        cudaEventRecord(e0);
        vlr_fp64_peak_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess || best > 1e29f) return VLR_ERR_CUDA;
    *tflops = 2.0 * 8.0 * (double)iters * (double)blocks * threads / ((double)best * 1e-3) / 1e12;
    return VLR_OK;
}

// ---- contamination estimator (contamination.cuh) ----
// Launch geometry of the likelihood kernel: events per CTA column, AFD points per shared-memory tile, CTA rows per SM.
// VLR_CONTAM_GEOM="threads,tile,rows_per_sm" overrides it for measurements (tile: 1024 or 2048).
struct ContamGeom {
    int threads, tile, rows_per_sm;
};
static ContamGeom contamination_geometry(int n_events) {
    ContamGeom g{416, 1024, 4}; // measured best of five (profiles/README.md): one column holds all 404 events
    if (const char* e = getenv("VLR_CONTAM_GEOM")) {
        int t = 0, tile = 0, r = 0;
        if (sscanf(e, "%d,%d,%d", &t, &tile, &r) == 3 && t >= 32 && t <= vlrcontam::CONTAM_MAX_THREADS && t % 32 == 0 &&
            (tile == 1024 || tile == 2048) && r >= 1 && r <= 32)
            g = ContamGeom{t, tile, r};
    }
    if (g.threads > ((n_events + 31) & ~31)) g.threads = (n_events + 31) & ~31;
    return g;
}

static vlr_status_t contamination_launch(int n_sms, const vlr_contamination_input_t* in, vlr_contamination_output_t* out,
                                         double* d_scratch, size_t scratch_doubles, cudaStream_t s) {
    using namespace vlrcontam;
    const int n_events = in->n_grid * in->n_max_vafs;
    const ContamGeom g = contamination_geometry(n_events);
    const int gx = (n_events + g.threads - 1) / g.threads;
    // chunks of observations: g.rows_per_sm CTA rows per SM, at least 32 observations each
    int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>((in->n_obs + 31) / 32, (int64_t)g.rows_per_sm * n_sms));
    const int64_t chunk = std::max<int64_t>(1, (in->n_obs + n_chunks - 1) / n_chunks);
    n_chunks = std::max<int64_t>(1, (in->n_obs + chunk - 1) / chunk);
    if ((size_t)(1 + n_chunks * n_events) > scratch_doubles) return VLR_ERR_INVALID_ARGUMENT;
    double* d_maxvaf = d_scratch;
    double* d_partial = d_scratch + 1;
    vlr_contam_maxvaf_kernel<<<1, 1024, 0, s>>>(in->max_posterior_vaf, in->n_obs, d_maxvaf, out->max_vaf);
    Obs o{in->n_obs, in->prob_denovo, in->max_posterior_vaf, in->afd_offsets, in->afd_vaf, in->afd_logp};
    const dim3 grid(gx, (unsigned)n_chunks);
    if (g.tile == 1024)
        vlr_contam_likelihood_kernel<1024><<<grid, g.threads, 0, s>>>(o, in->expected_max_somatic_vaf, in->n_max_vafs,
                                                                     in->n_grid, d_maxvaf, chunk, d_partial);
    else
        vlr_contam_likelihood_kernel<2048><<<grid, g.threads, 0, s>>>(o, in->expected_max_somatic_vaf, in->n_max_vafs,
                                                                     in->n_grid, d_maxvaf, chunk, d_partial);
    vlr_contam_finish_kernel<<<1, CONTAM_MAX_EVENTS, 0, s>>>(d_partial, (int)n_chunks, in->ln_prior, in->n_max_vafs,
                                                            in->n_grid, out->ln_posterior, out->ln_likelihood,
                                                            out->ln_marginal);
    return cudaGetLastError() == cudaSuccess ? VLR_OK : VLR_ERR_CUDA;
}

constexpr int CONTAM_MAX_ROWS_PER_SM = 32; // bound of ContamGeom::rows_per_sm, sizes the partial-sum scratch

static bool contamination_args_ok(const vlr_contamination_input_t* in, const vlr_contamination_output_t* out) {
    if (!in || !out || !out->ln_posterior || !out->ln_marginal) return false;
    if (in->n_obs < 0 || in->n_grid < 3 || in->n_grid % 2 == 0 || in->n_max_vafs < 1 ||
        in->n_max_vafs > vlrcontam::CONTAM_MAX_ROWS || in->n_grid * in->n_max_vafs > vlrcontam::CONTAM_MAX_EVENTS)
        return false;
    if (!in->expected_max_somatic_vaf || !in->ln_prior) return false;
    if (in->n_obs > 0 && (!in->prob_denovo || !in->max_posterior_vaf || !in->afd_offsets)) return false;
    return true;
}

static vlr_status_t contamination_device_ok(int32_t device, int* n_sms) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return VLR_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n_dev || cudaSetDevice(device) != cudaSuccess) return VLR_ERR_INVALID_ARGUMENT;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VLR_ERR_CUDA;
    if (prop.major < 10) return VLR_ERR_UNSUPPORTED;
    *n_sms = prop.multiProcessorCount;
    return VLR_OK;
}

vlr_status_t vlr_contamination_posterior_device(int32_t device, const vlr_contamination_input_t* in,
                                                vlr_contamination_output_t* out, void* cuda_stream) {
    if (!contamination_args_ok(in, out)) return VLR_ERR_INVALID_ARGUMENT;
    int n_sms = 0;
    vlr_status_t st = contamination_device_ok(device, &n_sms);
    if (st != VLR_OK) return st;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t scratch = 1 + (size_t)(CONTAM_MAX_ROWS_PER_SM * n_sms + 1) * (size_t)(in->n_grid * in->n_max_vafs);
    double* d_scratch = nullptr;
    if (cudaMallocAsync(&d_scratch, scratch * sizeof(double), s) != cudaSuccess) {
        cudaGetLastError();
        return VLR_ERR_OUT_OF_MEMORY;
    }
    st = contamination_launch(n_sms, in, out, d_scratch, scratch, s);
    cudaFreeAsync(d_scratch, s); // stream-ordered: released after the finishing kernel
    return st;
}

vlr_status_t vlr_contamination_gather_device(int32_t device, const vlr_results_t* results, int64_t n_loci,
                                             int32_t n_samples, int32_t n_events, int32_t sample, int32_t denovo_event,
                                             double min_prob, double* prob_denovo, double* max_posterior_vaf,
                                             int64_t* afd_offsets, double* afd_vaf, double* afd_logp, int64_t* kept_loci,
                                             int64_t* n_obs, void* cuda_stream) {
    using namespace vlrcontam;
    if (!results || !n_obs || n_loci < 0 || n_samples < 1 || n_events < 1 || sample < 0 || sample >= n_samples ||
        denovo_event < 0 || denovo_event >= n_events)
        return VLR_ERR_INVALID_ARGUMENT;
    if (results->afd_capacity < 1 || !results->afd_count || !results->afd_vaf || !results->afd_logp || !results->log_posteriors ||
        !results->map_vaf || !results->map_config || !results->status)
        return VLR_ERR_INVALID_ARGUMENT;
    if (!prob_denovo || !max_posterior_vaf || !afd_offsets || !afd_vaf || !afd_logp) return VLR_ERR_INVALID_ARGUMENT;
    int n_sms = 0;
    vlr_status_t st = contamination_device_ok(device, &n_sms);
    if (st != VLR_OK) return st;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    int64_t* d_tmp = nullptr; // obs_index [n], pt_offset [n], n_obs [1]
    if (cudaMallocAsync(&d_tmp, sizeof(int64_t) * (size_t)(2 * n_loci + 1), s) != cudaSuccess) {
        cudaGetLastError();
        return VLR_ERR_OUT_OF_MEMORY;
    }
    GatherArgs a;
    a.log_post = results->log_posteriors;
    a.map_vaf = results->map_vaf;
    a.map_config = results->map_config;
    a.status = results->status;
    a.afd_count = results->afd_count;
    a.afd_vaf = results->afd_vaf;
    a.afd_logp = results->afd_logp;
    a.n = n_loci;
    a.S = n_samples;
    a.E1 = n_events + 1;
    a.cap = results->afd_capacity;
    a.sample = sample;
    a.event = denovo_event;
    a.min_prob = min_prob;
    a.prob_denovo = prob_denovo;
    a.max_posterior_vaf = max_posterior_vaf;
    a.afd_offsets = afd_offsets;
    a.out_vaf = afd_vaf;
    a.out_logp = afd_logp;
    a.kept = kept_loci;
    a.obs_index = d_tmp;
    a.pt_offset = d_tmp + n_loci;
    a.n_obs = d_tmp + 2 * n_loci;
    vlr_contam_gather_scan_kernel<<<1, 1024, 0, s>>>(a);
    if (n_loci > 0) {
        const int blocks = (int)std::min<int64_t>((n_loci + 7) / 8, (int64_t)n_sms * 8);
        vlr_contam_gather_copy_kernel<<<blocks, 256, 0, s>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(n_obs, a.n_obs, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    cudaFreeAsync(d_tmp, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    return e == cudaSuccess ? VLR_OK : VLR_ERR_CUDA;
}

vlr_status_t vlr_contamination_posterior(int32_t device, const vlr_contamination_input_t* in,
                                         vlr_contamination_output_t* out) {
    if (!contamination_args_ok(in, out)) return VLR_ERR_INVALID_ARGUMENT;
    int n_sms = 0;
    vlr_status_t st = contamination_device_ok(device, &n_sms);
    if (st != VLR_OK) return st;
    const int64_t n = in->n_obs, n_pts = n > 0 ? in->afd_offsets[n] : 0;
    if (n_pts < 0 || (n_pts > 0 && (!in->afd_vaf || !in->afd_logp))) return VLR_ERR_INVALID_ARGUMENT;
    const int n_events = in->n_grid * in->n_max_vafs;
    const size_t scratch = 1 + (size_t)(CONTAM_MAX_ROWS_PER_SM * n_sms + 1) * (size_t)n_events;
    // one device block: [prob_denovo n][mpv n][afd_vaf P][afd_logp P][emsv R][prior G][post E][lik E][marginal][maxvaf]
    // [scratch] then the offsets (int64)
    const size_t n_d = 2 * (size_t)n + 2 * (size_t)n_pts + in->n_max_vafs + in->n_grid + 2 * (size_t)n_events + 2 + scratch;
    double* d = nullptr;
    int64_t* d_off = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t e = cudaSuccess;
    auto done = [&](vlr_status_t r) {
        if (s) cudaStreamDestroy(s);
        cudaFree(d);
        cudaFree(d_off);
        if (r != VLR_OK) cudaGetLastError();
        return r;
    };
    if (cudaMalloc(&d, n_d * sizeof(double)) != cudaSuccess || cudaMalloc(&d_off, ((size_t)n + 1) * sizeof(int64_t)) != cudaSuccess)
        return done(VLR_ERR_OUT_OF_MEMORY);
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return done(VLR_ERR_CUDA);
    double* p = d;
    auto put = [&](const double* src, size_t count) -> double* {
        double* at = p;
        p += count;
        if (count && e == cudaSuccess) e = cudaMemcpyAsync(at, src, count * sizeof(double), cudaMemcpyHostToDevice, s);
        return at;
    };
    vlr_contamination_input_t din = *in;
    din.prob_denovo = put(in->prob_denovo, (size_t)n);
    din.max_posterior_vaf = put(in->max_posterior_vaf, (size_t)n);
    din.afd_vaf = put(in->afd_vaf, (size_t)n_pts);
    din.afd_logp = put(in->afd_logp, (size_t)n_pts);
    din.expected_max_somatic_vaf = put(in->expected_max_somatic_vaf, (size_t)in->n_max_vafs);
    din.ln_prior = put(in->ln_prior, (size_t)in->n_grid);
    vlr_contamination_output_t dout;
    dout.ln_posterior = p;
    dout.ln_likelihood = p + n_events;
    dout.ln_marginal = p + 2 * n_events;
    dout.max_vaf = p + 2 * n_events + 1;
    double* d_scratch = p + 2 * n_events + 2;
    if (n > 0 && e == cudaSuccess)
        e = cudaMemcpyAsync(d_off, in->afd_offsets, ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    din.afd_offsets = d_off;
    if (e != cudaSuccess) return done(VLR_ERR_CUDA);
    st = contamination_launch(n_sms, &din, &dout, d_scratch, scratch, s);
    if (st != VLR_OK) return done(st);
    e = cudaMemcpyAsync(out->ln_posterior, dout.ln_posterior, n_events * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && out->ln_likelihood)
        e = cudaMemcpyAsync(out->ln_likelihood, dout.ln_likelihood, n_events * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out->ln_marginal, dout.ln_marginal, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && out->max_vaf)
        e = cudaMemcpyAsync(out->max_vaf, dout.max_vaf, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    return done(e == cudaSuccess ? VLR_OK : VLR_ERR_CUDA);
}

// NUMA node the current device hangs off (/sys/bus/pci/devices/<bus id>/numa_node), -1 when unknown.
static int device_numa_node() {
    int dev = 0;
    char bus[32] = {0};
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, (int)sizeof bus, dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    char path[96];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE* f = fopen(path, "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    return node;
}

// The pages of a pinned buffer are placed by the calling thread's memory policy when cudaHostAlloc faults them in. On a
// two-socket host with eight GPUs the default (the node the thread happens to run on) sends the copy engines of half
// the GPUs across the socket interconnect, which then bounds the host entries of ALL ranks of a job. Prefer the node
// of the current device for the duration of the allocation (MPOL_PREFERRED: falls back to other nodes when that one is
// full; a refused syscall - seccomp, no CAP_SYS_NICE - leaves the default placement). VLR_NUMA=0 switches it off.
void* vlr_host_alloc(size_t bytes) {
    void* p = nullptr;
    bool policy_set = false;
#ifdef __linux__
    int old_mode = 0;
    unsigned long old_mask[16] = {0};
    const char* env = getenv("VLR_NUMA");
    const int node = (env && env[0] == '0') ? -1 : device_numa_node();
    if (node >= 0 && node < (int)(sizeof old_mask * 8) &&
        syscall(SYS_get_mempolicy, &old_mode, old_mask, sizeof old_mask * 8, nullptr, 0UL) == 0) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] = 1UL << (node % (8 * sizeof(unsigned long)));
        policy_set = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8) == 0;
        if (getenv("VLR_NUMA_DEBUG")) fprintf(stderr, "vlr_host_alloc: %zu bytes, device node %d, policy %s\n", bytes, node, policy_set ? "set" : "refused");
    }
#endif
    const cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
#ifdef __linux__
    if (policy_set) syscall(SYS_set_mempolicy, old_mode, old_mode == 0 ? nullptr : old_mask, old_mode == 0 ? 0UL : sizeof old_mask * 8);
#endif
    if (e != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void vlr_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

vlr_status_t vlr_ctx_create(const vlr_scenario_t* scenario, int32_t device, vlr_ctx_t** out) {
    if (!out) return VLR_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!scenario) return VLR_ERR_INVALID_ARGUMENT;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return VLR_ERR_NO_DEVICE; // no CPU fallback by design
    }
    if (device < 0 || device >= n_dev) return VLR_ERR_INVALID_ARGUMENT;
    vlr_ctx* ctx = new vlr_ctx;
    auto bail = [&](vlr_status_t st) {
        // keep the message retrievable: callers get no ctx on failure, so print it
        fprintf(stderr, "vlr_ctx_create: %s\n", ctx->err.c_str());
        vlr_ctx_destroy(ctx);
        return st;
    };
    if (!ctx->prep.build(scenario)) {
        ctx->err = ctx->prep.error;
        return bail(VLR_ERR_INVALID_ARGUMENT);
    }
    ctx->device = device;
    ctx->S = scenario->n_samples;
    ctx->E = scenario->n_events;
    cudaError_t e;
#define CKB(call)                                            \
    if ((e = (call)) != cudaSuccess) {                       \
        ctx->fail_cuda(e, #call, __LINE__);                  \
        return bail(e == cudaErrorMemoryAllocation ? VLR_ERR_OUT_OF_MEMORY : VLR_ERR_CUDA); \
    }
    CKB(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKB(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        ctx->err = "device is not sm_100-class (this library is built for sm_100a only)";
        return bail(VLR_ERR_UNSUPPORTED);
    }
    ctx->n_sms = prop.multiProcessorCount;
    auto upload = [&](DevBuf& buf, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e2 = buf.ensure(bytes ? bytes : 8);
        if (e2 != cudaSuccess) return e2;
        if (bytes) e2 = cudaMemcpy(buf.p, src, bytes, cudaMemcpyHostToDevice);
        return e2;
    };
    CKB(upload(ctx->d_samples, scenario->samples, sizeof(vlr_sample_t) * scenario->n_samples));
    CKB(upload(ctx->d_events, scenario->events, sizeof(vlr_event_t) * scenario->n_events));
    CKB(upload(ctx->d_nodes, scenario->nodes, sizeof(vlr_node_t) * scenario->n_nodes));
    CKB(upload(ctx->d_set_vafs, scenario->set_vafs, sizeof(double) * (size_t)std::max(0, scenario->n_set_vafs)));
    CKB(upload(ctx->d_spectra, scenario->spectra, sizeof(vlr_spectrum_t) * (size_t)std::max(0, scenario->n_spectra)));
    CKB(upload(ctx->d_lfc_nodes, ctx->prep.lfc_nodes.data(), sizeof(int) * ctx->prep.lfc_nodes.size()));
    CKB(upload(ctx->d_lfc_ordinal, ctx->prep.lfc_ordinal.data(), sizeof(int) * ctx->prep.lfc_ordinal.size()));
    ctx->dsc = ctx->prep.view((const vlr_sample_t*)ctx->d_samples.p, (const vlr_event_t*)ctx->d_events.p,
                              (const vlr_node_t*)ctx->d_nodes.p, (const double*)ctx->d_set_vafs.p,
                              (const vlr_spectrum_t*)ctx->d_spectra.p, (const int*)ctx->d_lfc_nodes.p,
                              (const int*)ctx->d_lfc_ordinal.p);
    CKB(ctx->d_prior_tab.ensure(sizeof(PriorTabEntry) * PRIOR_TAB_N));
    CKB(cudaMemset(ctx->d_prior_tab.p, 0, sizeof(PriorTabEntry) * PRIOR_TAB_N));
    ctx->dsc.prior_tab = (PriorTabEntry*)ctx->d_prior_tab.p;
    // The tree walk recurses once per tree level (density -> subdensity -> density); frames hold an Ops copy and the
    // integration state. Size the per-thread stack from the scenario's depth.
    size_t stack = 4096 + (size_t)(ctx->prep.max_depth + 2) * 1024;
    CKB(cudaDeviceSetLimit(cudaLimitStackSize, stack));
    ctx->small = scenario->n_samples <= 3 && scenario->n_events <= 8 && ctx->prep.max_depth <= 6;
    ctx->ctx_stride = align16(ctx->small ? sizeof(vlr_small::Ctx) : sizeof(vlr_full::Ctx));
    ctx->smem_bytes = (size_t)WARPS_PER_CTA * ((size_t)ctx->ctx_stride + (size_t)SM_READS * 4 * sizeof(double));
    int per_sm = 0;
    if (ctx->small) {
        CKB(cudaFuncSetAttribute(vlr_call_kernel_vlr_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_bytes));
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vlr_call_kernel_vlr_small, THREADS, ctx->smem_bytes));
    } else {
        CKB(cudaFuncSetAttribute(vlr_call_kernel_vlr_full, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_bytes));
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vlr_call_kernel_vlr_full, THREADS, ctx->smem_bytes));
    }
    if (per_sm < 1) per_sm = 1;
    ctx->ctas_per_sm = per_sm;
    ctx->grid = per_sm * ctx->n_sms;
    ctx->wplan = ctx->prep.wave_plan();
    const char* wave_env = getenv("VLR_WAVE"); // VLR_WAVE=0: generic engine only (A/B measurements, debugging)
    ctx->wave = ctx->small && ctx->wplan.eligible && ctx->wplan.max_rounds <= vlr_small::W_MAXROUNDS &&
                !(wave_env && wave_env[0] == '0');
    if (ctx->wave) {
        ctx->wave_smem_prep = (size_t)WARPS_PER_CTA * (size_t)vlr_small::CTX_LEAN;
        CKB(cudaFuncSetAttribute(vlr_wave_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->wave_smem_prep));
        CKB(cudaFuncSetAttribute(vlr_wave_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->wave_smem_prep));
        CKB(cudaFuncSetAttribute(vlr_wave_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->wave_smem_prep));
        CKB(cudaFuncSetAttribute(vlr_wave_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WAVE_ROUND_SMEM));
        CKB(cudaFuncSetAttribute(vlr_wave_round_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WAVE_ROUND_SMEM));
        int n1 = 0, n2 = 0;
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, vlr_wave_coef_kernel, THREADS, ctx->wave_smem_prep));
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n2, vlr_wave_round_kernel, WAVE_ROUND_THREADS, WAVE_ROUND_SMEM));
        ctx->wave_grid_prep = std::max(1, n1) * ctx->n_sms;
        ctx->wave_grid_round = std::max(1, n2) * ctx->n_sms;
        // the finish kernel indexes the per-warp scratch (WarpWs) of the generic workspace: same number of warps or fewer
        int n3 = 0;
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n3, vlr_wave_finish_kernel, THREADS, ctx->wave_smem_prep));
        ctx->wave_grid_finish = std::min(std::max(1, n3) * ctx->n_sms, ctx->grid);
        const char* res_env = getenv("VLR_RESIDENT");
        ctx->resident = !(res_env && res_env[0] == '0');
        CKB(cudaFuncSetAttribute(vlr_wave_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem(8, vlr_small::R_SLOT_Q)));
        int n4 = 0;
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n4, vlr_wave_resident_kernel, RES_THREADS, res_smem(8, vlr_small::R_SLOT_Q)));
        if (getenv("VLR_WAVE_DEBUG")) fprintf(stderr, "vlr: class-1 resident kernel: %d CTAs per SM (%zu bytes of shared memory per CTA)\n", n4, res_smem(8, vlr_small::R_SLOT_Q));
        if (const char* e = getenv("VLR_RES_CTAS")) { // measurements: fewer CTAs of the octet kernel per SM
            const int v = atoi(e);
            if (v >= 1 && v < n4) n4 = v;
        }
        ctx->wave_grid_res = std::max(1, n4) * ctx->n_sms;
        CKB(cudaFuncSetAttribute(vlr_wave_resident_deep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem(32, vlr_small::R_SLOT_QL)));
        const int slots[3] = {vlr_small::R_SLOT_QM, vlr_small::R_SLOT_QD, vlr_small::R_SLOT_QL};
        for (int k = 0; k < 3; ++k) {
            int n5 = 0;
            CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n5, vlr_wave_resident_deep_kernel, RES_THREADS, res_smem(32, slots[k])));
            ctx->wave_grid_deep[k] = std::max(1, n5) * ctx->n_sms;
        }
    }
    {
        const char* sets_env = getenv("VLR_SETS"); // VLR_SETS=0: generic engine only (A/B measurements, tests)
        ctx->splan = ctx->prep.sets_plan();
        ctx->sets = ctx->small && !ctx->wave && ctx->splan.eligible && !(sets_env && sets_env[0] == '0');
    }
    if (ctx->sets) {
        SetsPlan& sp = ctx->splan;
        CKB(upload(ctx->d_sets_folds, ctx->prep.sets_folds.data(), sizeof(SetsFold) * ctx->prep.sets_folds.size()));
        CKB(upload(ctx->d_sets_leaves, ctx->prep.sets_leaves.data(), sizeof(SetsLeaf) * ctx->prep.sets_leaves.size()));
        CKB(upload(ctx->d_sets_leaf_vaf, ctx->prep.sets_leaf_vaf.data(), sizeof(double) * ctx->prep.sets_leaf_vaf.size()));
        CKB(ctx->d_sets_prior_val.ensure(sizeof(double) * 4 * (size_t)sp.n_leaves));
        CKB(ctx->d_sets_prior_side.ensure(sizeof(uint32_t) * 4 * (size_t)sp.n_leaves));
        CKB(ctx->d_sets_prior_state.ensure(sizeof(int) * 4));
        CKB(cudaMemset(ctx->d_sets_prior_state.p, 0, sizeof(int) * 4));
        sp.folds = (const SetsFold*)ctx->d_sets_folds.p;
        sp.leaves = (const SetsLeaf*)ctx->d_sets_leaves.p;
        sp.leaf_vaf = (const double*)ctx->d_sets_leaf_vaf.p;
        sp.prior_val = (double*)ctx->d_sets_prior_val.p;
        sp.prior_side = (uint32_t*)ctx->d_sets_prior_side.p;
        sp.prior_state = (int*)ctx->d_sets_prior_state.p;
        ctx->wave_smem_prep = (size_t)WARPS_PER_CTA * (size_t)vlr_small::CTX_LEAN; // (pre, finish: a lean Ctx per warp)
        ctx->sets_smem_lc = (size_t)WARPS_PER_CTA * SETS_LC_WARP_SMEM;
        CKB(cudaFuncSetAttribute(vlr_sets_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->wave_smem_prep));
        CKB(cudaFuncSetAttribute(vlr_sets_lc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->sets_smem_lc));
        CKB(cudaFuncSetAttribute(vlr_sets_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->wave_smem_prep));
        int n1 = 0, n2 = 0, n3 = 0;
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, vlr_sets_pre_kernel, THREADS, ctx->wave_smem_prep));
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n2, vlr_sets_lc_kernel, THREADS, ctx->sets_smem_lc));
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n3, vlr_sets_finish_kernel, THREADS, ctx->wave_smem_prep));
        ctx->sets_grid_pre = std::max(1, n1) * ctx->n_sms;
        ctx->sets_grid_lc = std::max(1, n2) * ctx->n_sms;
        // the finish kernel indexes the per-warp scratch (WarpWs) of the generic workspace: same number of warps or fewer
        ctx->sets_grid_finish = std::min(std::max(1, n3) * ctx->n_sms, ctx->grid);
    }
    CKB(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    if (const char* e = getenv("VLR_WAVE_STREAMS")) ctx->n_aux = std::max(1, std::min((int)vlr_ctx::MAX_AUX, atoi(e)));
    for (int i = 0; i < vlr_ctx::MAX_AUX; ++i) {
        CKB(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
        CKB(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
    }
    CKB(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CKB(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
    if (const char* e = getenv("VLR_NBUF")) {
        const int v = atoi(e);
        if (v >= 1 && v <= NBUF_MAX) ctx->nbuf = v;
    }
    for (int i = 0; i < NBUF_MAX; ++i) {
        CKB(cudaStreamCreateWithFlags(&ctx->slots[i].stream, cudaStreamNonBlocking));
        CKB(cudaEventCreateWithFlags(&ctx->slots[i].ev_ready, cudaEventDisableTiming));
        CKB(cudaEventCreateWithFlags(&ctx->slots[i].ev_free, cudaEventDisableTiming));
    }
    CKB(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
#undef CKB
    *out = ctx;
    return VLR_OK;
}

void vlr_ctx_destroy(vlr_ctx_t* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    auto free_slot = [](Slot& s) {
        if (s.stream) {
            cudaStreamSynchronize(s.stream);
            cudaStreamDestroy(s.stream);
        }
        if (s.ev_ready) cudaEventDestroy(s.ev_ready);
        if (s.ev_free) cudaEventDestroy(s.ev_free);
        s.offsets.release();
        for (auto& c : s.cols) c.release();
        for (auto& c : s.pk) c.release();
        for (auto& c : s.pk_dict) c.release();
        DevBuf* all[] = {&s.rflags, &s.hart, &s.hvar, &s.lflags, &s.het, &s.semr, &s.log_post, &s.log_marginal,
                         &s.map_vaf, &s.map_config, &s.best_event, &s.status, &s.n_base, &s.afd_count, &s.afd_vaf,
                         &s.afd_logp, &s.ws, &s.coef, &s.be, &s.ticket, &s.w_cnt, &s.w_loci, &s.w_lcs, &s.w_ogx,
                         &s.w_ogf, &s.w_coef, &s.w_tasks[0], &s.w_tasks[1], &s.w_list[0], &s.w_list[1], &s.w_dlist[0], &s.w_dlist[1], &s.w_deferred,
                         &s.w_gx, &s.w_gf, &s.w_be, &s.w_ben, &s.w_rlist, &s.w_rgx, &s.w_rgm, &s.w_rge, &s.w_rscratch,
                         &s.w_cscratch, &s.s_cnt, &s.s_loci, &s.s_lcs, &s.s_deferred, &s.s_be, &s.s_ben};
        for (DevBuf* b : all) b->release();
    };
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    free_slot(ctx->dev_slot);
    for (auto& sa : ctx->dev_slot_aux) free_slot(sa);
    for (int i = 0; i < vlr_ctx::MAX_AUX; ++i) {
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    for (auto& s : ctx->slots) free_slot(s);
    DevBuf* sc[] = {&ctx->d_samples, &ctx->d_events, &ctx->d_nodes, &ctx->d_set_vafs, &ctx->d_spectra,
                    &ctx->d_lfc_nodes, &ctx->d_lfc_ordinal, &ctx->d_prior_tab, &ctx->d_sets_folds, &ctx->d_sets_leaves,
                    &ctx->d_sets_leaf_vaf, &ctx->d_sets_prior_val, &ctx->d_sets_prior_side, &ctx->d_sets_prior_state};
    for (DevBuf* b : sc) b->release();
    delete ctx;
}

vlr_status_t vlr_ctx_reserve(vlr_ctx_t* ctx, int64_t max_reads_per_locus) {
    if (!ctx || max_reads_per_locus < 1) return VLR_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (max_reads_per_locus > ctx->reserve_reads) ctx->reserve_reads = max_reads_per_locus;
    return ensure_workspace(ctx, ctx->dev_slot, ctx->reserve_reads, false);
}

vlr_status_t vlr_call_batch_device(vlr_ctx_t* ctx, const vlr_batch_t* batch, vlr_results_t* results, void* cuda_stream) {
    if (!ctx) return VLR_ERR_INVALID_ARGUMENT;
    if (!batch_valid(batch) || !results_valid(results)) return ctx->fail(VLR_ERR_INVALID_ARGUMENT, "invalid batch or results");
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    vlr_status_t st = ensure_workspace(ctx, ctx->dev_slot, ctx->reserve_reads, results->afd_capacity > 0);
    if (st != VLR_OK) return st;
    if (ctx->have_done) CK(cudaStreamWaitEvent(stream, ctx->ev_done, 0)); // the workspace of the previous call is free
    DevBatch b;
    b.n_loci = batch->n_loci;
    b.read_base = 0;
    b.read_offsets = batch->read_offsets;
    b.pm = batch->prob_mapping;
    b.pr = batch->prob_ref;
    b.pa = batch->prob_alt;
    b.pmiss = batch->prob_missed_allele;
    b.psa = batch->prob_sample_alt;
    b.pdo = batch->prob_double_overlap;
    b.phb = batch->prob_hit_base;
    b.rflags = batch->read_flags;
    b.hart = batch->prob_homopolymer_artifact;
    b.hvar = batch->prob_homopolymer_variant;
    b.lflags = batch->locus_flags;
    b.het_phred = batch->locus_heterozygosity_phred;
    b.semr_phred = batch->locus_semr_phred;
    DevResults r;
    r.log_post = results->log_posteriors;
    r.log_marginal = results->log_marginal;
    r.map_vaf = results->map_vaf;
    r.map_config = results->map_config;
    r.best_event = results->best_event;
    r.status = results->status;
    r.n_base_events = results->n_base_events;
    r.afd_capacity = results->afd_capacity;
    r.afd_count = results->afd_count;
    r.afd_vaf = results->afd_vaf;
    r.afd_logp = results->afd_logp;
    const int64_t avg_reads = batch->n_loci > 0 ? (batch->n_reads + batch->n_loci - 1) / batch->n_loci : 0;
    // a batch is split over the internal streams from 131 072 loci or ~100 M reads on (100 000 config-5 loci = 150 M
    // reads: one stream 1.43, three 1.51 M loci/s)
    int64_t split_min = batch->n_reads >= 3 * ((int64_t)1 << 25) ? 1 << 14 : 1 << 17;
    if (const char* e = getenv("VLR_WAVE_SPLIT_MIN")) split_min = std::max<long>(1024, atol(e)); // tuning knob (measurements)
    if (ctx->wave && ctx->n_aux > 1 && batch->n_loci >= split_min) {
        // large batch on the wavefront pipeline: its parts run on internal streams (own workspaces), forked from and
        // joined to the caller's stream, so that one part's kernels fill the grid tails and the straggler rounds of
        // the others (the host entry gets the same effect from its three chunk streams)
        const int ns = ctx->n_aux;
        CK(cudaEventRecord(ctx->ev_fork, stream));
        for (int i = 0; i < ns; ++i) {
            Slot& sl = i == 0 ? ctx->dev_slot : ctx->dev_slot_aux[i - 1];
            if (i > 0) {
                st = ensure_workspace(ctx, sl, ctx->reserve_reads, results->afd_capacity > 0);
                if (st != VLR_OK) return st;
            }
            const int64_t lo = batch->n_loci * i / ns, hi = batch->n_loci * (i + 1) / ns;
            CK(cudaStreamWaitEvent(ctx->aux[i], ctx->ev_fork, 0));
            st = launch_wave(ctx, sl, b, r, avg_reads, ctx->aux[i], lo, hi);
            if (st != VLR_OK) return st;
            CK(cudaEventRecord(ctx->ev_join[i], ctx->aux[i]));
            CK(cudaStreamWaitEvent(stream, ctx->ev_join[i], 0));
        }
        CK(cudaEventRecord(ctx->ev_done, stream));
        ctx->have_done = true;
        return VLR_OK;
    }
    st = launch(ctx, ctx->dev_slot, b, r, stream, avg_reads);
    if (st != VLR_OK) return st;
    CK(cudaEventRecord(ctx->ev_done, stream));
    ctx->have_done = true;
    return VLR_OK;
}

static vlr_status_t call_batch_host(vlr_ctx_t* ctx, const vlr_packed_batch_t* batch, vlr_results_t* results) {
    CK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    int64_t unpack_launches = 0;
    const int S = ctx->S, E = ctx->E;
    const int nbuf = ctx->nbuf;
    const int64_t L = batch->n_loci;
    if (L == 0) return VLR_OK;
    const int64_t* off = batch->read_offsets;
    if (off[L * S] != batch->n_reads) return ctx->fail(VLR_ERR_INVALID_ARGUMENT, "read_offsets[n_loci*S] != n_reads");
    // chunking: a chunk must hold many more loci than the 2368 resident warps (dynamic load balance inside the kernel)
    // and tens of MB of columns (PCIe efficiency); 3 slots of <= 16M reads (512 MB) stay far below the HBM size
    int64_t target_reads = 16 << 20;
    int64_t max_loci_chunk = 1 << 16;
    if (const char* e = getenv("VLR_CHUNK_LOCI")) { // tuning knob (measurements): loci and reads per chunk scale together
        const long v = atol(e);
        if (v >= 1024 && v <= (1 << 22)) {
            max_loci_chunk = v;
            target_reads = std::max<int64_t>(target_reads, v * 256);
        }
    }
    {
        // equal chunks, a multiple of the nbuf streams of them: the streams finish together (a greedy cut leaves the last
        // chunks alone on the device)
        int64_t n_chunks = std::max((L + max_loci_chunk - 1) / max_loci_chunk, (batch->n_reads + target_reads - 1) / target_reads);
        if (n_chunks > 1) n_chunks = (n_chunks + nbuf - 1) / nbuf * nbuf;
        n_chunks = std::max<int64_t>(n_chunks, 1);
        max_loci_chunk = std::min(max_loci_chunk, (L + n_chunks - 1) / n_chunks);
        target_reads = std::min(target_reads, batch->n_reads / n_chunks + batch->n_reads / (n_chunks * 64) + 1);
    }
    const int cap = results->afd_capacity;
    static const size_t enc_bytes[5] = {4, 2, 2, 1, 0}; // VLR_ENC_*: bytes per read on the link
    bool any_packed = false;
    for (int c = 0; c < VLR_N_PACKED_COLUMNS; ++c) any_packed = any_packed || batch->columns[c].encoding != VLR_ENC_F32;
    int64_t lo = 0;
    int k = 0;
    vlr_status_t st = VLR_OK;
    const bool timing = getenv("VLR_CHUNK_TIMING") != nullptr; // measurements: per-chunk timeline on stderr
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s) {
        if (!timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tev.push_back(e);
    };
    while (lo < L) {
        int64_t hi = lo, max_reads = 0;
        const int64_t r0 = off[lo * S];
        while (hi < L && hi - lo < max_loci_chunk) {
            int64_t n = off[(hi + 1) * S] - off[hi * S];
            if (n < 0) return ctx->fail(VLR_ERR_INVALID_ARGUMENT, "read_offsets must be non-decreasing");
            if (hi > lo && off[(hi + 1) * S] - r0 > target_reads) break;
            if (n > max_reads) max_reads = n;
            ++hi;
        }
        const int64_t r1 = off[hi * S], nl = hi - lo, nr = r1 - r0;
        Slot& sl = ctx->slots[k % nbuf];
        // nbuf slots of buffers, nbuf / 2 compute streams, ONE copy stream: the copies of chunk k start as soon as the
        // kernels of chunk k - nbuf have read the slot's inputs (ev_free), i.e. while the chunks in front of it are still
        // computing, and its kernels (on the stream of chunk k - nbuf / 2) start when the copies are done (ev_ready).
        // The host queues the whole batch without waiting; a buffer that has to grow is freed first, which synchronises.
        const int ncomp = std::max(1, nbuf / 2);
        cudaStream_t s = ctx->slots[(k % nbuf) % ncomp].stream, hs = ctx->copy_stream; // a slot always computes on the same stream
        st = ensure_workspace(ctx, sl, max_reads, cap > 0);
        if (st != VLR_OK) break;
        if (k >= nbuf) CK(cudaStreamWaitEvent(hs, sl.ev_free, 0));
        mark(hs);
        CK(sl.offsets.ensure(sizeof(int64_t) * (size_t)(nl * S + 1)));
        CK(cudaMemcpyAsync(sl.offsets.p, off + lo * S, sizeof(int64_t) * (size_t)(nl * S + 1), cudaMemcpyHostToDevice, hs));
        // columns: plain ones go straight into place, encoded ones into the slot's staging buffers (with their
        // dictionaries the first time the slot is used in this call) and are widened on the device
        UnpackParams up;
        up.n = nr;
        for (int c = 0; c < VLR_N_PACKED_COLUMNS; ++c) {
            const vlr_column_t& col = batch->columns[c];
            DevBuf& dstb = c == VLR_COL_READ_FLAGS ? sl.rflags : sl.cols[c];
            CK(dstb.ensure(sizeof(uint32_t) * (size_t)std::max<int64_t>(nr, 4)));
            up.enc[c] = col.encoding;
            up.n_dict[c] = col.n_dict;
            up.dst[c] = (uint32_t*)dstb.p;
            up.src[c] = nullptr;
            up.dict[c] = nullptr;
            const size_t w = enc_bytes[col.encoding];
            if (col.encoding == VLR_ENC_F32) {
                if (nr) CK(cudaMemcpyAsync(dstb.p, (const char*)col.data + 4 * (size_t)r0, 4 * (size_t)nr, cudaMemcpyHostToDevice, hs));
                continue;
            }
            if (w) {
                CK(sl.pk[c].ensure(w * (size_t)std::max<int64_t>(nr, 4)));
                if (nr) CK(cudaMemcpyAsync(sl.pk[c].p, (const char*)col.data + w * (size_t)r0, w * (size_t)nr, cudaMemcpyHostToDevice, hs));
                up.src[c] = sl.pk[c].p;
            }
            if (col.n_dict > 0) {
                if (k < nbuf) {
                    CK(sl.pk_dict[c].ensure(sizeof(uint32_t) * (size_t)col.n_dict));
                    CK(cudaMemcpyAsync(sl.pk_dict[c].p, col.dict, sizeof(uint32_t) * (size_t)col.n_dict, cudaMemcpyHostToDevice, hs));
                }
                up.dict[c] = (const uint32_t*)sl.pk_dict[c].p;
            }
        }
        auto opt_up = [&](DevBuf& buf, const float* src, int64_t first, int64_t n) -> cudaError_t {
            if (!src) return cudaSuccess;
            cudaError_t e = buf.ensure(sizeof(float) * (size_t)std::max<int64_t>(n, 1));
            if (e != cudaSuccess || n == 0) return e;
            return cudaMemcpyAsync(buf.p, src + first, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, hs);
        };
        CK(opt_up(sl.hart, batch->prob_homopolymer_artifact, r0, nr));
        CK(opt_up(sl.hvar, batch->prob_homopolymer_variant, r0, nr));
        CK(opt_up(sl.het, batch->locus_heterozygosity_phred, lo, nl));
        CK(opt_up(sl.semr, batch->locus_semr_phred, lo, nl));
        CK(sl.lflags.ensure(sizeof(uint32_t) * (size_t)nl));
        CK(cudaMemcpyAsync(sl.lflags.p, batch->locus_flags + lo, sizeof(uint32_t) * (size_t)nl, cudaMemcpyHostToDevice, hs));
        CK(cudaEventRecord(sl.ev_ready, hs));
        mark(hs);
        CK(cudaStreamWaitEvent(s, sl.ev_ready, 0));
        if (any_packed && nr) {
            const int bx = (int)std::min<int64_t>((nr / 4 + 255) / 256 + 1, (int64_t)ctx->n_sms * 8);
            vlr_unpack_kernel<<<dim3(bx, VLR_N_PACKED_COLUMNS), 256, 0, s>>>(up);
            CK(cudaGetLastError());
            unpack_launches++;
        }
        // results
        CK(sl.log_post.ensure(sizeof(double) * (size_t)(nl * (E + 1))));
        CK(sl.log_marginal.ensure(sizeof(double) * (size_t)nl));
        CK(sl.map_vaf.ensure(sizeof(double) * (size_t)(nl * S)));
        CK(sl.map_config.ensure(sizeof(int32_t) * (size_t)nl));
        CK(sl.best_event.ensure(sizeof(int32_t) * (size_t)nl));
        CK(sl.status.ensure(sizeof(uint32_t) * (size_t)nl));
        CK(sl.n_base.ensure(sizeof(uint32_t) * (size_t)nl));
        if (cap > 0) {
            CK(sl.afd_count.ensure(sizeof(int32_t) * (size_t)(nl * S)));
            CK(sl.afd_vaf.ensure(sizeof(double) * (size_t)(nl * S * cap)));
            CK(sl.afd_logp.ensure(sizeof(double) * (size_t)(nl * S * cap)));
        }
        DevBatch b;
        b.n_loci = nl;
        b.read_base = r0; // offsets stay absolute; columns hold rows [r0, r1)
        b.read_offsets = (const int64_t*)sl.offsets.p;
        b.pm = (const float*)sl.cols[0].p;
        b.pr = (const float*)sl.cols[1].p;
        b.pa = (const float*)sl.cols[2].p;
        b.pmiss = (const float*)sl.cols[3].p;
        b.psa = (const float*)sl.cols[4].p;
        b.pdo = (const float*)sl.cols[5].p;
        b.phb = (const float*)sl.cols[6].p;
        b.rflags = (const uint32_t*)sl.rflags.p;
        b.hart = batch->prob_homopolymer_artifact ? (const float*)sl.hart.p : nullptr;
        b.hvar = batch->prob_homopolymer_variant ? (const float*)sl.hvar.p : nullptr;
        b.lflags = (const uint32_t*)sl.lflags.p;
        b.het_phred = batch->locus_heterozygosity_phred ? (const float*)sl.het.p : nullptr;
        b.semr_phred = batch->locus_semr_phred ? (const float*)sl.semr.p : nullptr;
        DevResults r;
        r.log_post = (double*)sl.log_post.p;
        r.log_marginal = (double*)sl.log_marginal.p;
        r.map_vaf = (double*)sl.map_vaf.p;
        r.map_config = (int32_t*)sl.map_config.p;
        r.best_event = (int32_t*)sl.best_event.p;
        r.status = (uint32_t*)sl.status.p;
        r.n_base_events = (uint32_t*)sl.n_base.p;
        r.afd_capacity = cap;
        r.afd_count = (int32_t*)sl.afd_count.p;
        r.afd_vaf = (double*)sl.afd_vaf.p;
        r.afd_logp = (double*)sl.afd_logp.p;
        mark(s);
        st = launch(ctx, sl, b, r, s, nl > 0 ? (nr + nl - 1) / nl : 0);
        if (st != VLR_OK) break;
        CK(cudaEventRecord(sl.ev_free, s));
        mark(s);
        // D2H
        CK(cudaMemcpyAsync(results->log_posteriors + lo * (E + 1), r.log_post, sizeof(double) * (size_t)(nl * (E + 1)), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(results->map_vaf + lo * S, r.map_vaf, sizeof(double) * (size_t)(nl * S), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(results->status + lo, r.status, sizeof(uint32_t) * (size_t)nl, cudaMemcpyDeviceToHost, s));
        if (results->log_marginal) CK(cudaMemcpyAsync(results->log_marginal + lo, r.log_marginal, sizeof(double) * (size_t)nl, cudaMemcpyDeviceToHost, s));
        if (results->map_config) CK(cudaMemcpyAsync(results->map_config + lo, r.map_config, sizeof(int32_t) * (size_t)nl, cudaMemcpyDeviceToHost, s));
        if (results->best_event) CK(cudaMemcpyAsync(results->best_event + lo, r.best_event, sizeof(int32_t) * (size_t)nl, cudaMemcpyDeviceToHost, s));
        if (results->n_base_events) CK(cudaMemcpyAsync(results->n_base_events + lo, r.n_base_events, sizeof(uint32_t) * (size_t)nl, cudaMemcpyDeviceToHost, s));
        if (cap > 0) {
            CK(cudaMemcpyAsync(results->afd_count + lo * S, r.afd_count, sizeof(int32_t) * (size_t)(nl * S), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(results->afd_vaf + lo * S * cap, r.afd_vaf, sizeof(double) * (size_t)(nl * S * cap), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(results->afd_logp + lo * S * cap, r.afd_logp, sizeof(double) * (size_t)(nl * S * cap), cudaMemcpyDeviceToHost, s));
        }
        mark(s);
        lo = hi;
        ++k;
    }
    ctx->launches += unpack_launches;
    for (int i = 0; i <= nbuf; ++i) {
        cudaError_t e = cudaStreamSynchronize(i < nbuf ? ctx->slots[i].stream : ctx->copy_stream);
        if (e != cudaSuccess && st == VLR_OK) st = ctx->fail_cuda(e, "cudaStreamSynchronize", __LINE__);
    }
    if (timing && !tev.empty()) {
        for (size_t c = 0; c + 4 < tev.size(); c += 5) {
            float t[5];
            for (int j = 0; j < 5; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[c + j]);
            fprintf(stderr, "chunk %2zu slot %zu: h2d %7.2f .. %7.2f  kernels %7.2f .. %7.2f  d2h .. %7.2f ms\n", c / 5, (c / 5) % (size_t)nbuf, t[0], t[1], t[2], t[3], t[4]);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return st;
}


static bool packed_valid(const vlr_packed_batch_t* b) {
    if (!b || b->n_loci < 0 || b->n_reads < 0) return false;
    if (b->n_loci == 0) return true;
    if (!b->read_offsets || !b->locus_flags) return false;
    for (int c = 0; c < VLR_N_PACKED_COLUMNS; ++c) {
        const vlr_column_t& col = b->columns[c];
        switch (col.encoding) {
        case VLR_ENC_F32: if (b->n_reads > 0 && !col.data) return false; break;
        case VLR_ENC_F16: if (c == VLR_COL_READ_FLAGS || (b->n_reads > 0 && !col.data)) return false; break;
        case VLR_ENC_DICT16: if (col.n_dict < 1 || col.n_dict > 65536 || !col.dict || (b->n_reads > 0 && !col.data)) return false; break;
        case VLR_ENC_DICT8: if (col.n_dict < 1 || col.n_dict > 256 || !col.dict || (b->n_reads > 0 && !col.data)) return false; break;
        case VLR_ENC_CONST: if (col.n_dict != 1 || !col.dict) return false; break;
        default: return false;
        }
    }
    return true;
}

vlr_status_t vlr_call_batch(vlr_ctx_t* ctx, const vlr_batch_t* batch, vlr_results_t* results) {
    if (!ctx) return VLR_ERR_INVALID_ARGUMENT;
    if (!batch_valid(batch) || !results_valid(results)) return ctx->fail(VLR_ERR_INVALID_ARGUMENT, "invalid batch or results");
    vlr_packed_batch_t pk; // the plain batch as a packed one whose columns are all F32: one code path
    memset(&pk, 0, sizeof pk);
    pk.n_loci = batch->n_loci;
    pk.n_reads = batch->n_reads;
    pk.read_offsets = batch->read_offsets;
    const void* cols[VLR_N_PACKED_COLUMNS] = {batch->prob_mapping, batch->prob_ref, batch->prob_alt, batch->prob_missed_allele,
                                              batch->prob_sample_alt, batch->prob_double_overlap, batch->prob_hit_base, batch->read_flags};
    for (int c = 0; c < VLR_N_PACKED_COLUMNS; ++c) {
        pk.columns[c].encoding = VLR_ENC_F32;
        pk.columns[c].data = cols[c];
    }
    pk.prob_homopolymer_artifact = batch->prob_homopolymer_artifact;
    pk.prob_homopolymer_variant = batch->prob_homopolymer_variant;
    pk.locus_flags = batch->locus_flags;
    pk.locus_heterozygosity_phred = batch->locus_heterozygosity_phred;
    pk.locus_semr_phred = batch->locus_semr_phred;
    return call_batch_host(ctx, &pk, results);
}

// ---- vlr_pack_batch: host-side encoder of the packed batch -----------------------------------------------------------
} // extern "C"
namespace {

struct PackedOwner { // what vlr_pack_batch hands out: the public struct first, then what it owns
    vlr_packed_batch_t pub;
    std::vector<std::pair<void*, bool>> bufs; // (pointer, page-locked?)
};

// open addressing over 32-bit patterns; slot value = pattern + 1 as 64 bits (0 = empty)
struct PatternSet {
    static constexpr uint32_t CAP = 1u << 18, MASK = CAP - 1;
    std::vector<uint64_t> slot;
    std::vector<uint32_t> code; // same index as slot: dictionary position (filled by number())
    uint32_t n = 0;
    PatternSet() : slot(CAP, 0) {}
    static uint32_t hash(uint32_t v) {
        v *= 0x9E3779B1u;
        return (v ^ (v >> 15)) & MASK;
    }
    // false: more than 65536 distinct patterns (no dictionary)
    bool insert(uint32_t v) {
        uint32_t h = hash(v);
        const uint64_t key = (uint64_t)v + 1;
        for (;;) {
            const uint64_t s = slot[h];
            if (s == key) return true;
            if (s == 0) {
                if (n >= 65536) return false;
                slot[h] = key;
                ++n;
                return true;
            }
            h = (h + 1) & MASK;
        }
    }
    uint32_t find(uint32_t v) const {
        uint32_t h = hash(v);
        const uint64_t key = (uint64_t)v + 1;
        while (slot[h] != key) h = (h + 1) & MASK;
        return code[h];
    }
    std::vector<uint32_t> values() const {
        std::vector<uint32_t> out;
        out.reserve(n);
        for (uint64_t s : slot)
            if (s) out.push_back((uint32_t)(s - 1));
        return out;
    }
    void number(const std::vector<uint32_t>& sorted) {
        code.assign(CAP, 0);
        for (uint32_t i = 0; i < (uint32_t)sorted.size(); ++i) {
            uint32_t h = hash(sorted[i]);
            const uint64_t key = (uint64_t)sorted[i] + 1;
            while (slot[h] != key) h = (h + 1) & MASK;
            code[h] = i;
        }
    }
};

// the f32 pattern is exactly representable as an IEEE half (normal, subnormal, zero, inf; NaNs are left to F32/dictionary)
inline bool half_exact(uint32_t b) {
    const uint32_t e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
    if (e == 0xffu) return m == 0;           // inf
    if (e == 0) return m == 0;               // zero (f32 subnormals are below the half range)
    const int ue = (int)e - 127;
    if (ue > 15) return false;
    if (ue >= -14) return (m & 0x1fffu) == 0; // normal half: 10 mantissa bits
    if (ue < -24) return false;
    const int drop = 13 + (-14 - ue);         // subnormal half: fewer bits
    return (m & ((1u << drop) - 1u)) == 0;
}
inline uint16_t to_half_exact(uint32_t b) { // only for patterns half_exact() accepted
    const uint32_t sgn = (b >> 16) & 0x8000u, e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
    if (e == 0xffu) return (uint16_t)(sgn | 0x7c00u);
    if (e == 0) return (uint16_t)sgn;
    const int ue = (int)e - 127;
    if (ue >= -14) return (uint16_t)(sgn | ((uint32_t)(ue + 15) << 10) | (m >> 13));
    const uint32_t full = m | 0x800000u;
    return (uint16_t)(sgn | (full >> (13 + (-14 - ue))));
}

template <class F>
void parallel_ranges(int64_t n, int threads, F f) {
    if (threads <= 1 || n < (1 << 16)) {
        f(0, (int64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(f, t, n * t / threads, n * (t + 1) / threads);
    for (auto& x : th) x.join();
}

} // namespace
extern "C" {

vlr_status_t vlr_pack_batch(const vlr_batch_t* batch, int32_t n_threads, vlr_packed_batch_t** out) {
    if (!out || !batch_valid(batch)) return VLR_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    PackedOwner* own = new (std::nothrow) PackedOwner();
    if (!own) return VLR_ERR_OUT_OF_MEMORY;
    vlr_packed_batch_t& pk = own->pub;
    memset(&pk, 0, sizeof pk);
    pk.n_loci = batch->n_loci;
    pk.n_reads = batch->n_reads;
    pk.read_offsets = batch->read_offsets;
    pk.prob_homopolymer_artifact = batch->prob_homopolymer_artifact;
    pk.prob_homopolymer_variant = batch->prob_homopolymer_variant;
    pk.locus_flags = batch->locus_flags;
    pk.locus_heterozygosity_phred = batch->locus_heterozygosity_phred;
    pk.locus_semr_phred = batch->locus_semr_phred;
    const uint32_t* cols[VLR_N_PACKED_COLUMNS] = {
        (const uint32_t*)batch->prob_mapping,    (const uint32_t*)batch->prob_ref,
        (const uint32_t*)batch->prob_alt,        (const uint32_t*)batch->prob_missed_allele,
        (const uint32_t*)batch->prob_sample_alt, (const uint32_t*)batch->prob_double_overlap,
        (const uint32_t*)batch->prob_hit_base,   batch->read_flags};
    const int64_t n = batch->n_reads;
    auto pinned = [&](size_t bytes) -> void* { // page-locked when a device is there (plain memory still works, staged)
        void* p = vlr_host_alloc(bytes ? bytes : 1);
        const bool locked = p != nullptr;
        if (!p) p = malloc(bytes ? bytes : 1);
        if (p) own->bufs.emplace_back(p, locked);
        return p;
    };
    vlr_status_t st = VLR_OK;
    for (int c = 0; c < VLR_N_PACKED_COLUMNS && st == VLR_OK; ++c) {
        vlr_column_t& col = pk.columns[c];
        const uint32_t* src = cols[c];
        col.encoding = VLR_ENC_F32;
        col.data = src;
        if (n == 0) continue;
        // pass 1: distinct patterns per thread (one-entry memo: neighbouring reads repeat), half-exactness
        std::vector<PatternSet> sets((size_t)threads);
        std::vector<char> overflow((size_t)threads, 0), all_half((size_t)threads, 1);
        parallel_ranges(n, threads, [&](int t, int64_t lo, int64_t hi) {
            PatternSet& ps = sets[(size_t)t];
            bool ok = true, half = c != VLR_COL_READ_FLAGS;
            uint32_t last = ~src[lo];
            for (int64_t i = lo; i < hi && (ok || half); ++i) {
                const uint32_t v = src[i];
                if (v == last) continue;
                last = v;
                if (ok && !ps.insert(v)) ok = false;
                if (half && !half_exact(v)) half = false;
            }
            overflow[(size_t)t] = ok ? 0 : 1;
            all_half[(size_t)t] = half ? 1 : 0;
        });
        bool dict_ok = true, half_ok = c != VLR_COL_READ_FLAGS;
        for (int t = 0; t < threads; ++t) {
            dict_ok = dict_ok && !overflow[(size_t)t];
            half_ok = half_ok && all_half[(size_t)t];
        }
        PatternSet& all = sets[0];
        for (int t = 1; t < threads && dict_ok; ++t)
            for (uint32_t v : sets[(size_t)t].values())
                if (!all.insert(v)) {
                    dict_ok = false;
                    break;
                }
        std::vector<uint32_t> dict;
        if (dict_ok) {
            dict = all.values();
            std::sort(dict.begin(), dict.end());
            all.number(dict);
        }
        int enc = VLR_ENC_F32;
        if (dict_ok && dict.size() == 1) enc = VLR_ENC_CONST;
        else if (dict_ok && dict.size() <= 256) enc = VLR_ENC_DICT8;
        else if (half_ok) enc = VLR_ENC_F16;
        else if (dict_ok) enc = VLR_ENC_DICT16;
        if (enc == VLR_ENC_F32) continue;
        col.encoding = enc;
        if (enc != VLR_ENC_F16) {
            uint32_t* d = (uint32_t*)pinned(sizeof(uint32_t) * dict.size());
            if (!d) {
                st = VLR_ERR_OUT_OF_MEMORY;
                break;
            }
            memcpy(d, dict.data(), sizeof(uint32_t) * dict.size());
            col.dict = d;
            col.n_dict = (int32_t)dict.size();
        }
        if (enc == VLR_ENC_CONST) {
            col.data = nullptr;
            continue;
        }
        const size_t w = enc == VLR_ENC_DICT8 ? 1 : 2;
        void* data = pinned(w * (size_t)n);
        if (!data) {
            st = VLR_ERR_OUT_OF_MEMORY;
            break;
        }
        col.data = data;
        // pass 2: codes
        parallel_ranges(n, threads, [&](int, int64_t lo, int64_t hi) {
            uint32_t last = ~src[lo], code = 0;
            for (int64_t i = lo; i < hi; ++i) {
                const uint32_t v = src[i];
                if (v != last) {
                    last = v;
                    code = enc == VLR_ENC_F16 ? (uint32_t)to_half_exact(v) : all.find(v);
                }
                if (w == 1) ((uint8_t*)data)[i] = (uint8_t)code;
                else ((uint16_t*)data)[i] = (uint16_t)code;
            }
        });
    }
    if (st != VLR_OK) {
        vlr_packed_batch_free(&own->pub);
        return st;
    }
    *out = &own->pub;
    return VLR_OK;
}

void vlr_packed_batch_free(vlr_packed_batch_t* packed) {
    if (!packed) return;
    PackedOwner* own = reinterpret_cast<PackedOwner*>(packed); // `pub` is the first member
    for (auto& b : own->bufs) {
        if (b.second) vlr_host_free(b.first);
        else free(b.first);
    }
    delete own;
}

int64_t vlr_packed_batch_bytes(const vlr_packed_batch_t* pk, int32_t n_samples) {
    if (!pk) return 0;
    static const int64_t w[5] = {4, 2, 2, 1, 0};
    int64_t bytes = 8 * (pk->n_loci * n_samples + 1) + 4 * pk->n_loci;
    for (int c = 0; c < VLR_N_PACKED_COLUMNS; ++c) {
        const vlr_column_t& col = pk->columns[c];
        if (col.encoding < 0 || col.encoding > 4) return -1;
        bytes += w[col.encoding] * pk->n_reads + 4 * (int64_t)col.n_dict;
    }
    if (pk->prob_homopolymer_artifact) bytes += 4 * pk->n_reads;
    if (pk->prob_homopolymer_variant) bytes += 4 * pk->n_reads;
    if (pk->locus_heterozygosity_phred) bytes += 4 * pk->n_loci;
    if (pk->locus_semr_phred) bytes += 4 * pk->n_loci;
    return bytes;
}

vlr_status_t vlr_call_batch_packed(vlr_ctx_t* ctx, const vlr_packed_batch_t* packed, vlr_results_t* results) {
    if (!ctx) return VLR_ERR_INVALID_ARGUMENT;
    if (!packed_valid(packed) || !results_valid(results)) return ctx->fail(VLR_ERR_INVALID_ARGUMENT, "invalid packed batch or results");
    return call_batch_host(ctx, packed, results);
}

} // extern "C"
