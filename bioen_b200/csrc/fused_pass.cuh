// fused_pass.cuh -- the forces method in TWO passes over the matrix (the algorithmic minimum)
// instead of the reference's five (c_bioen_kernels_forces.c:43-76) or the four of the unfused tile kernels.
//
// The forces evaluation alternates a reduction over the observables i (per structure j) with a reduction over
// the structures j (per observable i):
//
//   F1   x_j = sum_i f_i y_ij  ->  e_j = w0_j exp(x_j - max)  ->  avg_i ~ sum_j y_ij e_j       kernels_forces.c:128-171, 93-109
//   F2   t_j = sum_i r_i y_ij  ->  E_j = (theta (1 + lr_j) + t_j) w_j  ->  grad_i = sum_j (y_ij - avg_i) E_j      :298-338
//
// In the reference's observable-major layout (M x N) the first reduction needs a whole COLUMN before the second
// can start, and a column block of useful width does not fit in shared memory.  So the forces method keeps a
// second, structure-major copy  Yt[j][i]  (N x ldt; the reference's own `yTildeT` cache, c_bioen.pyx:391-393,
// made once on the device by k_transpose): all M observables of a structure are contiguous, a slab of C whole
// structures is ONE 1-D bulk copy (cp.async.bulk, completion on an mbarrier), and both reductions run on the
// same copy of the data -- see fused_team_pass below.  HBM traffic: N*ldt*8 bytes per kernel, read exactly once.
// All reductions are fixed-order: bit-reproducible run to run.
//
// History (measured at N = 1e6, M = 1e3, 8 GB per pass; 1.15 ms = the unfused tile kernel = HBM roofline):
//   block-wide kernel, 8 warps per slab, two barriers per slab            2.35 ms   latency chain per slab
//   teams + producer warp with 64-bit divisions in its loop               1.85 ms   producer-bound
//   teams + division-free producer warp                                   1.34 ms   producer-bound (8 KB copies)
//   teams feeding their own rings, inputs riding with the slab (this)     1.22 ms
// F1 (softmax) runs ~10 % slower than F2 (gradient) although it does less arithmetic: timing builds without the
// exp, without the x_j store and without any softmax logic all stayed at 1.21-1.23 ms (F2: 1.10-1.12 ms), and the
// ncu source view shows F1's warps waiting on the slab barrier (28 % of samples) where F2's never do.
#pragma once
#include <float.h>

#include "common.cuh"

namespace bioen {

constexpr int kFCMax = 8;                   // structures per slab
constexpr int kFMinM = 256;                 // below this the thread mapping is mostly idle: use the tile kernels
constexpr int kFMaxLdt = 8192;              // 16 pair-slots x 64 lanes x 8 warps

enum FusedKind { kFusedSoftmaxAvg = 0, kFusedGradient = 1 };

// ------------------------------------------------------------------------------------------------
// Team variant (ldt <= 4096): the block-wide kernel above synchronises all 8 warps twice per slab and its
// per-slab dependency chain (shuffle reductions, exp) is longer than the time the slab takes to arrive, so it
// runs at half the HBM rate.  Here a TEAM of T warps (T = 1 for M <= 1024) owns whole structures on its own:
// the structure's M values are loaded from shared memory ONCE into registers, dotted with a_i, reduced with
// shuffles (plus one T-warp named barrier when T > 1), turned into the per-structure factor, and accumulated
// into the team's register-resident M-vector.  The 8/T teams of a CTA run on different structures with no
// synchronisation between them, so one team's latency chain hides behind the others' arithmetic.  Each team
// has its own ring of slabs (its own full/empty mbarriers); one producer lane feeds all rings.
// ------------------------------------------------------------------------------------------------
constexpr int kTWarps = 8;                   // consumer warps per CTA
constexpr int kTThreads = kTWarps * 32;      // no producer warp: every team feeds its own ring
constexpr int kTMaxRing = 32;                // teams * stages
constexpr int kTMaxLdt = 8192;               // T = 8: one team of all eight warps
constexpr int kTAuxBytes = 256;              // per ring slot: two windows of <= 10 doubles (w0 | w, lr)

struct TeamArgs {
    const double* Yt;
    long long ldt;
    int M, N;
    int C;                 // structures per slab (<= kFCMax)
    int stages;            // ring depth per team
    long long nslab, chunk;
    const double* ab;      // interleaved {a_i, *}
    const double* b;       // F2: avg_i
    const double* s0;      // w0_j (F1) / w_j (F2)
    const double* s1;      // lr_j (F2)
    double theta;
    double* xout;
    double* part;          // [gridDim.x][ldp]  (the teams of a CTA are combined in the kernel)
    long long ldp;
    double* lse;           // [gridDim.x][2]
    int evict_first;       // 1: the matrix is much larger than L2, stream it through
};

template <int KI, int T, int KIND>
__global__ void __launch_bounds__(kTThreads, 1) fused_team_pass(const TeamArgs a) {
    constexpr int TEAMS = kTWarps / T;
    // NOTE: index the extern array directly.  Rounding the base through uintptr_t makes the compiler lose the
    // shared address space and emit generic LD.E instead of LDS.  Bulk copies need 16-byte alignment only,
    // which the declaration guarantees.
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int slab_bytes = a.C * (int)a.ldt * 8;
    const int aux_off = (slab_bytes + 127) & ~127;       // per-structure inputs ride along with the slab
    const int stage_bytes = aux_off + kTAuxBytes;
    unsigned char* ring = smem_raw;                                          // [TEAMS][stages][stage_bytes]
    double* a_s = reinterpret_cast<double*>(ring + (size_t)TEAMS * a.stages * stage_bytes);   // [ldt]
    double* b_s = a_s + a.ldt;                                               // [ldt] (F2 only)
    uint64_t* full = reinterpret_cast<uint64_t*>(b_s + (KIND == kFusedGradient ? a.ldt : 0));
    double* tred = reinterpret_cast<double*>(full + kTMaxRing);              // [TEAMS][2][T]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long s_begin = (long long)blockIdx.x * a.chunk;
    const long long s_end = (s_begin + a.chunk < a.nslab) ? s_begin + a.chunk : a.nslab;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TEAMS * a.stages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < a.ldt; i += blockDim.x) {
        a_s[i] = a.ab[2 * i];
        if (KIND == kFusedGradient) b_s[i] = a.b[i];
    }
    __syncthreads();

    // There is no producer warp: a single lane cannot issue one 8 KB bulk copy every ~300 cycles (measured:
    // the kernel ran producer-bound).  Every team feeds its own ring: as soon as a slab has been read into
    // registers its slot is refilled with the slab `stages` rounds ahead, so all slots but the one being read
    // are always in flight and no empty-barrier handshake is needed.
    const int team = warp / T, wt = warp % T;
    const bool issuer = (wt == 0 && lane == 0);
    const uint64_t policy = a.evict_first ? l2_policy_evict_first() : l2_policy_evict_last();
    auto issue = [&](long long s, int slot) {
        const long long j0 = s * a.C;
        const int Cs = (a.N - j0 < a.C) ? (int)(a.N - j0) : a.C;
        const uint32_t bytes = (uint32_t)Cs * (uint32_t)a.ldt * 8u;
        // per-structure inputs (w0_j, or w_j and lr_j): 16-byte aligned window around [j0, j0 + Cs).  They must
        // not be fetched by dependent loads: with ~20 MB of bulk copies queued on the chip a DRAM load that
        // the arithmetic waits for costs several microseconds.
        const long long jb = j0 & ~1LL;
        const uint32_t abytes = (uint32_t)(((j0 + Cs - jb) + 1) & ~1LL) * 8u;
        unsigned char* st = ring + (size_t)slot * stage_bytes;
        mbar_expect_tx(&full[slot], bytes + abytes * (KIND == kFusedGradient ? 2u : 1u));
        bulk_load_1d_hint(st, a.Yt + (size_t)j0 * a.ldt, bytes, &full[slot], policy);
        bulk_load_1d(st + aux_off, a.s0 + jb, abytes, &full[slot]);
        if (KIND == kFusedGradient) bulk_load_1d(st + aux_off + kTAuxBytes / 2, a.s1 + jb, abytes, &full[slot]);
    };
    if (issuer) {
        for (int i = 0; i < a.stages; ++i) {
            const long long s = s_begin + team + (long long)TEAMS * i;
            if (s < s_end) issue(s, team * a.stages + i);
        }
    }

    const int p0 = 2 * (lane + 32 * wt);   // this thread's observable pairs: p0 + 64*T*k
    double2 acc[KI];
#pragma unroll
    for (int k = 0; k < KI; ++k) acc[k] = make_double2(0.0, 0.0);
    double m_run = -DBL_MAX, S_run = 0.0;
    int par = 0;
    int stage = 0;
    uint32_t phase = 0;
    const long long refill = (long long)TEAMS * a.stages;

    for (long long s = s_begin + team; s < s_end; s += TEAMS) {
        const int slot = team * a.stages + stage;
        const long long j0 = s * a.C;
        const int Cs = (a.N - j0 < a.C) ? (int)(a.N - j0) : a.C;
        mbar_wait(&full[slot], phase);
        const double* slab = reinterpret_cast<const double*>(ring + (size_t)slot * stage_bytes);
        const double* aux0 = reinterpret_cast<const double*>(ring + (size_t)slot * stage_bytes + aux_off) + (j0 & 1);
        const double* aux1 = aux0 + kTAuxBytes / 16;
#pragma unroll 1
        for (int c = 0; c < Cs; ++c) {
            const double q0 = aux0[c];
            const double q1 = (KIND == kFusedGradient) ? aux1[c] : 0.0;
            // ---- phase A: the structure's values into registers, dot with a_i
            double2 y[KI];
            double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
            for (int k = 0; k < KI; ++k) {
                y[k] = make_double2(0.0, 0.0);
                if ((p0 + 64 * T * k) < a.ldt) {
                    y[k] = *reinterpret_cast<const double2*>(slab + (size_t)c * a.ldt + (p0 + 64 * T * k));
                    const double2 av = *reinterpret_cast<const double2*>(a_s + (p0 + 64 * T * k));
                    if (k & 1) { d2 = fma(y[k].x, av.x, d2); d3 = fma(y[k].y, av.y, d3); }
                    else       { d0 = fma(y[k].x, av.x, d0); d1 = fma(y[k].y, av.y, d1); }
                }
            }
            double x = warp_sum((d0 + d1) + (d2 + d3));
            if (T > 1) {
                double* tr = tred + (team * 2 + par) * T;
                if (lane == 0) tr[wt] = x;
                asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(T * 32) : "memory");
                x = 0.0;
#pragma unroll
                for (int w = 0; w < T; ++w) x += tr[w];
                par ^= 1;
            }
            if (c == Cs - 1) {
                // the whole slab has been read (this structure is in registers, q0/q1 too; for T > 1 the team
                // barrier above ordered every warp's reads): refill the slot for `stages` rounds ahead
                __syncwarp();
                if (issuer && s + refill < s_end) {
                    fence_proxy_async();
                    issue(s + refill, slot);
                }
            }
            // ---- per-structure factor
            double v;
            if (KIND == kFusedSoftmaxAvg) {
                if (x > m_run) {   // new running maximum: rescale what has been accumulated (rare)
                    const double sc = exp(m_run - x);
                    S_run *= sc;
#pragma unroll
                    for (int k = 0; k < KI; ++k) { acc[k].x *= sc; acc[k].y *= sc; }
                    m_run = x;
                }
                v = q0 * exp(x - m_run);
                S_run += v;
                if (wt == 0 && lane == 0) a.xout[j0 + c] = x;
            } else {
                v = ((1.0 + q1) * a.theta + x) * q0;
            }
            // ---- phase B: accumulate from registers
#pragma unroll
            for (int k = 0; k < KI; ++k) {
                if (KIND == kFusedGradient) {
                    if ((p0 + 64 * T * k) < a.ldt) {
                        const double2 bv = *reinterpret_cast<const double2*>(b_s + (p0 + 64 * T * k));
                        acc[k].x = fma(y[k].x - bv.x, v, acc[k].x);
                        acc[k].y = fma(y[k].y - bv.y, v, acc[k].y);
                    }
                } else {
                    acc[k].x = fma(y[k].x, v, acc[k].x);
                    acc[k].y = fma(y[k].y, v, acc[k].y);
                }
            }
        }
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }

    // ---- combine the teams of the CTA in shared memory (the ring is free now: every copy issued was consumed)
    __syncthreads();
    double* tm = reinterpret_cast<double*>(ring);                 // [TEAMS][ldt]
    double* tm_m = tred;                                          // [TEAMS]
    double* tm_S = tred + TEAMS;                                  // [TEAMS]  (tred holds 2*kTWarps doubles)
#pragma unroll
    for (int k = 0; k < KI; ++k) {
        if ((p0 + 64 * T * k) < a.ldt)
            *reinterpret_cast<double2*>(tm + (size_t)team * a.ldt + (p0 + 64 * T * k)) = acc[k];
    }
    if (KIND == kFusedSoftmaxAvg && wt == 0 && lane == 0) {
        tm_m[team] = m_run;
        tm_S[team] = S_run;
    }
    __syncthreads();
    double scale[TEAMS];
    double m_cta = -DBL_MAX, S_cta = 0.0;
    if (KIND == kFusedSoftmaxAvg) {
#pragma unroll
        for (int t = 0; t < TEAMS; ++t) m_cta = fmax(m_cta, tm_m[t]);
#pragma unroll
        for (int t = 0; t < TEAMS; ++t) {
            scale[t] = exp(tm_m[t] - m_cta);
            S_cta = fma(tm_S[t], scale[t], S_cta);
        }
    } else {
#pragma unroll
        for (int t = 0; t < TEAMS; ++t) scale[t] = 1.0;
    }
    double* out = a.part + (size_t)blockIdx.x * a.ldp;
    for (int i = threadIdx.x; i < a.ldt; i += blockDim.x) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < TEAMS; ++t) v = fma(tm[(size_t)t * a.ldt + i], scale[t], v);
        out[i] = v;
    }
    if (KIND == kFusedSoftmaxAvg && threadIdx.x == 0) {
        a.lse[2 * blockIdx.x] = m_cta;
        a.lse[2 * blockIdx.x + 1] = S_cta;
    }
}

// merge the rows' running (max, sum) pairs into this rank's pair (one block, fixed-order tree)
__global__ void __launch_bounds__(256) k_fused_lse_merge(int nrows, const double* lse, double* sc_pair) {
    __shared__ double sm[256], ss[256];
    double m = -DBL_MAX, s = 0.0;
    for (int c = threadIdx.x; c < nrows; c += 256) {
        const double m2 = lse[2 * c], s2 = lse[2 * c + 1];
        const double mm = fmax(m, m2);
        s = s * exp(m - mm) + s2 * exp(m2 - mm);
        m = mm;
    }
    sm[threadIdx.x] = m;
    ss[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double m1 = sm[threadIdx.x], m2 = sm[threadIdx.x + o];
            const double mm = fmax(m1, m2);
            ss[threadIdx.x] = ss[threadIdx.x] * exp(m1 - mm) + ss[threadIdx.x + o] * exp(m2 - mm);
            sm[threadIdx.x] = mm;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sc_pair[0] = sm[0];
        sc_pair[1] = ss[0];
    }
}

// out_i = sum_rows part[row][i] * scale_row;  scale_row = exp(m_row - M) / S with (M, S) the global pair (F1), or 1.
// One block per 32 observables; 8 row-groups per block split the rows, combined in a fixed order.
__global__ void __launch_bounds__(256) k_fused_merge_rows(int m, int nrows, const double* part, long long ldp,
                                                          const double* lse_rows, const double* lse_pairs, int nranks,
                                                          double* out) {
    __shared__ double red[8][33];
    double M = 0.0, inv = 1.0;
    if (lse_rows) {
        double S;
        M = lse_pairs[0]; S = lse_pairs[1];
        for (int r = 1; r < nranks; ++r) {
            const double m2 = lse_pairs[2 * r], s2 = lse_pairs[2 * r + 1];
            const double mm = fmax(M, m2);
            S = S * exp(M - mm) + s2 * exp(m2 - mm);
            M = mm;
        }
        inv = 1.0 / S;
    }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    double s = 0.0;
    for (int c = ty; c < nrows; c += 8) {
        const double scale = lse_rows ? exp(lse_rows[2 * c] - M) * inv : 1.0;
        if (i < m) s = fma(part[(size_t)c * ldp + i], scale, s);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < m) {
        double t = red[0][tx];
#pragma unroll
        for (int g = 1; g < 8; ++g) t += red[g][tx];
        out[i] = t;
    }
}

// Yt[j][i] = Y[i][j]  (32 x 32 tiles through shared memory); pad column (ldt > M) is zeroed
__global__ void __launch_bounds__(256) k_transpose(const double* __restrict__ Y, long long ld, int M, int N,
                                                   double* __restrict__ Yt, long long ldt) {
    __shared__ double tile[32][33];
    const long long j0 = (long long)blockIdx.x * 32;
    const int i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r;
        const long long j = j0 + tx;
        tile[r][tx] = (i < M && j < N) ? Y[(size_t)i * ld + j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const long long j = j0 + r;
        const int i = i0 + tx;
        if (j < N && i < ldt) Yt[(size_t)j * ldt + i] = tile[tx][r];
    }
}

// Yt[j][row0 + r] = src[r][j] for a block of `nr` rows that arrived row-major (structure-major-only uploads)
__global__ void __launch_bounds__(256) k_transpose_block(const double* __restrict__ src, long long lds, int nr, int N,
                                                         double* __restrict__ Yt, long long ldt, int row0) {
    __shared__ double tile[32][33];
    const long long j0 = (long long)blockIdx.x * 32;
    const int r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = r0 + r;
        const long long j = j0 + tx;
        tile[r][tx] = (i < nr && j < N) ? src[(size_t)i * lds + j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const long long j = j0 + r;
        const int i = r0 + tx;
        if (j < N && i < nr) Yt[(size_t)j * ldt + row0 + i] = tile[tx][r];
    }
}

// out[r][c] = Yt[col0 + c][row0 + r]  (downloads from the structure-major copy)
__global__ void __launch_bounds__(256) k_gather_block(const double* __restrict__ Yt, long long ldt, int row0, int nr,
                                                      long long col0, long long nc, double* __restrict__ out,
                                                      long long ldo) {
    __shared__ double tile[32][33];
    const long long c0 = (long long)blockIdx.x * 32;
    const int r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = ty; q < 32; q += 8) {
        const long long c = c0 + q;
        const int r = r0 + tx;
        tile[q][tx] = (c < nc && r < nr) ? Yt[(size_t)(col0 + c) * ldt + row0 + r] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = ty; q < 32; q += 8) {
        const int r = r0 + q;
        const long long c = c0 + tx;
        if (r < nr && c < nc) out[(size_t)r * ldo + c] = tile[tx][q];
    }
}

}  // namespace bioen
