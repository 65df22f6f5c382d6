// contamination.cuh — the contamination estimator's second Bayesian model (src/estimation/contamination.rs) on the GPU.
//
// The reference collects, for every call with P(denovo) >= 0.95, a VariantObservation {prob_denovo, the sample's
// allele frequency distribution, its MAP allele frequency} (contamination.rs:44-82) and then evaluates, for each of
// 4 x 101 events (expected maximum somatic VAF x contamination), prior + sum over observations of the AFD density at
// the VAF the event predicts for the observation (contamination.rs:163-186), integrates each of the 4 rows with an
// 101-point Simpson rule and normalises (contamination.rs:213-240). Work = observations x 404 interpolations.
//
// Mapping: one THREAD per event, one CTA row (blockIdx.y) per contiguous chunk of observations. A CTA stages the AFDs
// of a tile of its observations in shared memory with coalesced loads; every thread then walks the tile's observations
// sequentially and adds its event's term in a register, so there is no reduction inside the CTA and the result does
// not depend on the launch geometry of blockIdx.x. Partial sums per (chunk, event) are combined in chunk order by the
// finishing kernel (deterministic), which also runs the Simpson rows and the normalisation.
#pragma once

#include "engine_types.cuh"

namespace vlrcontam {
using namespace vlrcore;

struct Obs { // device or host view of the CSR-packed VariantObservations
    int64_t n_obs;
    const double* prob_denovo;       // [n_obs] ln P(denovo)
    const double* max_posterior_vaf; // [n_obs]
    const int64_t* afd_offsets;      // [n_obs + 1]
    const double* afd_vaf;           // ascending within an observation (the reference's BTreeMap order)
    const double* afd_logp;
};

// VariantObservation::pdf (contamination.rs:84-115) on one observation's sorted AFD.
VLR_DEV double obs_pdf(const double* vaf, const double* logp, int n, double x) {
    int lo = 0, hi = n; // first index with vaf >= x  (`vaf_dist.range(vaf..).next()`)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (vaf[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    if (lo < n && vaf[lo] == x) return logp[lo]; // case 1: exact
    if (lo == 0 || lo == n) return neg_inf();    // cases 3-5: outside the support / empty
    // case 2: linear interpolation of the density, written in log space exactly like the reference. A falling
    // segment makes the inner log's argument negative: NaN there too (`f64::ln`), and it propagates.
    const double inf_d = logp[lo - 1], sup_d = logp[lo];
    const double slope = (m_exp(sup_d) - m_exp(inf_d)) / (vaf[lo] - vaf[lo - 1]);
    return ln_add_exp(inf_d, m_log(slope) + m_log(x - vaf[lo - 1]));
}

// contamination of grid point i: itertools-num linspace(0, 1, n) as rust-bio's Simpson rule walks it
VLR_DEV double grid_contamination(int i, int n_grid) { return 0.0 + (double)i * ((1.0 - 0.0) / (double)(n_grid - 1)); }

// Likelihood::compute's per-observation term (contamination.rs:167-184)
VLR_DEV double obs_term(double prob_denovo, double mpv, double max_vaf, double emsv, double purity, const double* vaf,
                        const double* logp, int n) {
    if (purity == 0.0) return ln_one_minus_exp(prob_denovo); // there cannot be any denovo somatic mutation
    const double quantile = mpv / max_vaf;                   // VAFDist::get_expected_vaf (contamination.rs:263-271)
    return obs_pdf(vaf, logp, n, emsv * purity * quantile);
}

constexpr int CONTAM_MAX_THREADS = 512; // events per CTA column (blockDim.x is chosen by the launcher)
constexpr int CONTAM_TILE_OBS = 128;

#ifndef VLR_HOST_EMU
// VAFDist::new's max_vaf (contamination.rs:249-258): the largest MAP allele frequency, 0.0 without observations.
// One CTA; `copy` (optional) receives the value for the caller.
__global__ void __launch_bounds__(1024)
vlr_contam_maxvaf_kernel(const double* __restrict__ mpv, int64_t n_obs, double* __restrict__ max_vaf,
                         double* __restrict__ copy) {
    __shared__ double s_max[32];
    double m = 0.0;
    for (int64_t o = threadIdx.x; o < n_obs; o += blockDim.x) m = fmax(m, mpv[o]);
    for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0.0;
        for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (threadIdx.x == 0) {
            *max_vaf = m;
            if (copy) *copy = m;
        }
    }
}

// grid = (ceil(n_events / blockDim.x), n_chunks); partial[chunk][event]. CONTAM_TILE_PTS = AFD points staged per tile
// (16 B each of shared memory).
template <int CONTAM_TILE_PTS>
__global__ void __launch_bounds__(CONTAM_MAX_THREADS)
vlr_contam_likelihood_kernel(Obs in, const double* __restrict__ emsv, int n_emsv, int n_grid,
                             const double* __restrict__ max_vaf_p, int64_t chunk, double* __restrict__ partial) {
    const int CONTAM_THREADS = blockDim.x;
    __shared__ double s_vaf[CONTAM_TILE_PTS], s_lp[CONTAM_TILE_PTS];
    __shared__ double s_pd[CONTAM_TILE_OBS], s_mpv[CONTAM_TILE_OBS];
    __shared__ int s_off[CONTAM_TILE_OBS + 1];
    const int n_events = n_emsv * n_grid;
    const int e = blockIdx.x * CONTAM_THREADS + threadIdx.x;
    const bool live = e < n_events;
    const int k = live ? e / n_grid : 0, gi = live ? e % n_grid : 0;
    const double my_emsv = emsv[k], max_vaf = *max_vaf_p;
    const double purity = 1.0 - grid_contamination(gi, n_grid);
    const int64_t o_end = min(in.n_obs, ((int64_t)blockIdx.y + 1) * chunk);
    int64_t o = (int64_t)blockIdx.y * chunk;
    double sum = 0.0; // LogProb's Sum starts from ln_one = 0.0
    while (o < o_end) {
        // tile = the longest run [o, oe) of at most CONTAM_TILE_OBS observations whose AFDs fit the staging buffers
        // (uniform binary search over the offsets); an observation that alone exceeds them is read from global memory
        const int64_t base = in.afd_offsets[o];
        int64_t lo = o, hi = min(o_end, o + CONTAM_TILE_OBS);
        while (lo < hi) {
            int64_t mid = (lo + hi + 1) >> 1;
            if (in.afd_offsets[mid] - base <= CONTAM_TILE_PTS) lo = mid;
            else hi = mid - 1;
        }
        const int64_t oe = lo;
        if (oe == o) {
            if (live) {
                const int n = (int)(in.afd_offsets[o + 1] - base);
                sum += obs_term(in.prob_denovo[o], in.max_posterior_vaf[o], max_vaf, my_emsv, purity, in.afd_vaf + base,
                                in.afd_logp + base, n);
            }
            o += 1;
            continue;
        }
        const int n_tile = (int)(oe - o), n_pts = (int)(in.afd_offsets[oe] - base);
        for (int j = threadIdx.x; j < n_pts; j += CONTAM_THREADS) {
            s_vaf[j] = in.afd_vaf[base + j];
            s_lp[j] = in.afd_logp[base + j];
        }
        for (int j = threadIdx.x; j <= n_tile; j += CONTAM_THREADS) {
            s_off[j] = (int)(in.afd_offsets[o + j] - base);
            if (j < n_tile) {
                s_pd[j] = in.prob_denovo[o + j];
                s_mpv[j] = in.max_posterior_vaf[o + j];
            }
        }
        __syncthreads();
        if (live)
            for (int j = 0; j < n_tile; ++j)
                sum += obs_term(s_pd[j], s_mpv[j], max_vaf, my_emsv, purity, s_vaf + s_off[j], s_lp + s_off[j],
                                s_off[j + 1] - s_off[j]);
        __syncthreads();
        o = oe;
    }
    if (live) partial[(int64_t)blockIdx.y * n_events + e] = sum;
}
#endif

// ln_sum_exp in rust-bio's two-pass form (first maximum, its index skipped, terms added in slice order)
VLR_DEV double ln_sum_exp_slice(const double* p, int n) {
    if (n == 0) return neg_inf();
    double pmax = p[0];
    int imax = 0;
    for (int i = 1; i < n; ++i)
        if (p[i] > pmax) {
            pmax = p[i];
            imax = i;
        }
    if (pmax == neg_inf()) return neg_inf();
    if (pmax == INFINITY) return pmax;
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        if (i == imax || p[i] == neg_inf()) continue;
        s += m_exp(p[i] - pmax);
    }
    return pmax + m_log1p(s);
}

// One row of Marginal::compute (contamination.rs:222-236): ln_simpsons_integrate_exp(density, 0, 1, n_grid) over the
// row's joints; `scratch` holds n_grid values. Interior points first (weights 4, 2, 4, ...), then f(a), f(b).
VLR_DEV double simpson_row(const double* joint, int n_grid, double* scratch) {
    int m = 0;
    for (int i = 1; i < n_grid - 1; ++i) scratch[m++] = joint[i] + m_log((double)(2 + (i % 2) * 2));
    scratch[m++] = joint[0];
    scratch[m++] = joint[n_grid - 1];
    return ln_sum_exp_slice(scratch, m) + m_log(1.0 - 0.0) - m_log((double)(n_grid - 1)) - m_log(3.0);
}

constexpr int CONTAM_MAX_EVENTS = 1024;
constexpr int CONTAM_MAX_ROWS = 8;

#ifndef VLR_HOST_EMU
// One CTA of CONTAM_MAX_EVENTS threads: joint = prior + sum of the chunk partials (in chunk order), Simpson per row,
// marginal, posteriors. ln_posterior[k][i] = joint - marginal (ModelInstance::event_posteriors).
__global__ void __launch_bounds__(CONTAM_MAX_EVENTS)
vlr_contam_finish_kernel(const double* __restrict__ partial, int n_chunks, const double* __restrict__ ln_prior,
                         int n_emsv, int n_grid, double* __restrict__ ln_posterior, double* __restrict__ ln_likelihood,
                         double* __restrict__ ln_marginal) {
    __shared__ double s_joint[CONTAM_MAX_EVENTS];
    __shared__ double s_scratch[CONTAM_MAX_EVENTS];
    __shared__ double s_rows[CONTAM_MAX_ROWS];
    __shared__ double s_marginal;
    const int n_events = n_emsv * n_grid, e = threadIdx.x;
    if (e < n_events) {
        double lik = 0.0;
        for (int c = 0; c < n_chunks; ++c) lik += partial[(int64_t)c * n_events + e];
        if (ln_likelihood) ln_likelihood[e] = lik;
        s_joint[e] = ln_prior[e % n_grid] + lik;
    }
    __syncthreads();
    if (e < n_emsv) s_rows[e] = simpson_row(s_joint + e * n_grid, n_grid, s_scratch + e * n_grid);
    __syncthreads();
    if (e == 0) {
        s_marginal = ln_sum_exp_slice(s_rows, n_emsv);
        *ln_marginal = s_marginal;
    }
    __syncthreads();
    if (e < n_events) ln_posterior[e] = s_joint[e] - s_marginal;
}
#endif


// ---- VariantObservation::new for a batch of calls, on the device (contamination.rs:44-82) ----------------------------
// The estimator is a CallProcessor on `call_generic` (contamination.rs:282-306): per call it keeps {P(denovo), the
// sample's allele frequency distribution, its MAP allele frequency} when P(denovo) >= 0.95 and an AFD exists. With
// device-resident results of vlr_call_batch_device the same selection and the CSR packing of the kept AFDs run on the
// device: a single CTA counts and scans (calls x 2 ints: tiny), a warp per kept call copies.
struct GatherArgs {
    const double* log_post;    // [n][E + 1]
    const double* map_vaf;     // [n][S]
    const int32_t* map_config; // [n]
    const uint32_t* status;    // [n]
    const int32_t* afd_count;  // [n][S]
    const double* afd_vaf;     // [n][S][cap]
    const double* afd_logp;
    int64_t n;
    int S, E1, cap, sample, event;
    double min_prob;
    // out
    double* prob_denovo;
    double* max_posterior_vaf;
    int64_t* afd_offsets;
    double* out_vaf;
    double* out_logp;
    int64_t* kept; // optional: call index of every observation
    int64_t* n_obs;
    // scratch
    int64_t* obs_index; // [n]: observation number of the call, or -1
    int64_t* pt_offset; // [n]
};

VLR_DEV bool gather_keeps(const GatherArgs& a, int64_t l) {
    if (a.status[l] & VLR_ST_NO_MAP) return false;                 // sample_infos returned None
    if (a.map_config[l] != 0) return false;                       // artifact MAP: no allele frequency distribution
    return m_exp(a.log_post[l * a.E1 + a.event]) >= a.min_prob;   // prob_denovo.exp() >= 0.95
}

#ifndef VLR_HOST_EMU
__global__ void __launch_bounds__(1024) vlr_contam_gather_scan_kernel(const GatherArgs a) {
    __shared__ long long s_obs[1024], s_pts[1024];
    const int t = (int)threadIdx.x, T = (int)blockDim.x;
    const int64_t chunk = (a.n + T - 1) / T, lo = min(a.n, (int64_t)t * chunk), hi = min(a.n, lo + chunk);
    long long n_obs = 0, n_pts = 0;
    for (int64_t l = lo; l < hi; ++l)
        if (gather_keeps(a, l)) {
            n_obs++;
            n_pts += a.afd_count[l * a.S + a.sample];
        }
    s_obs[t] = n_obs;
    s_pts[t] = n_pts;
    __syncthreads();
    for (int o = 1; o < T; o <<= 1) { // inclusive scan (Hillis-Steele: 1024 entries)
        const long long vo = t >= o ? s_obs[t - o] : 0, vp = t >= o ? s_pts[t - o] : 0;
        __syncthreads();
        s_obs[t] += vo;
        s_pts[t] += vp;
        __syncthreads();
    }
    long long obs = s_obs[t] - n_obs, pts = s_pts[t] - n_pts; // exclusive
    for (int64_t l = lo; l < hi; ++l) {
        if (gather_keeps(a, l)) {
            a.obs_index[l] = obs;
            a.pt_offset[l] = pts;
            obs++;
            pts += a.afd_count[l * a.S + a.sample];
        } else {
            a.obs_index[l] = -1;
        }
    }
    if (t == T - 1) {
        *a.n_obs = s_obs[t];
        a.afd_offsets[s_obs[t]] = s_pts[t];
    }
}

__global__ void __launch_bounds__(256) vlr_contam_gather_copy_kernel(const GatherArgs a) {
    const int lane = (int)(threadIdx.x & 31);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t l = warp; l < a.n; l += n_warps) {
        const int64_t o = a.obs_index[l];
        if (o < 0) continue;
        const int64_t at = a.pt_offset[l];
        const int cnt = a.afd_count[l * a.S + a.sample];
        if (lane == 0) {
            a.prob_denovo[o] = a.log_post[l * a.E1 + a.event];
            a.max_posterior_vaf[o] = a.map_vaf[l * a.S + a.sample];
            a.afd_offsets[o] = at;
            if (a.kept) a.kept[o] = l;
        }
        const int64_t src = (l * a.S + a.sample) * a.cap;
        for (int k = lane; k < cnt; k += 32) {
            a.out_vaf[at + k] = a.afd_vaf[src + k];
            a.out_logp[at + k] = a.afd_logp[src + k];
        }
    }
}

#endif // VLR_HOST_EMU

} // namespace vlrcontam
