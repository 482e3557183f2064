// particles.cu — particle-side kernels and the step orchestration:
//   k_evolve (evolve_kernel.cuh), radix sort by (cell, tet), segmented per-cell
//   moment reduction, particle-number control, injection, seeding, I/O.
// Compiled with -fmad=false (see evolve_kernel.cuh).
//
// Reference: mcParticleCloud::evolve (mcParticle/mcParticleCloud/mcParticleCloud.C:1226-1530),
// updateCloudPDF first half (:929-956), particleNumberControl (:1014-1221),
// initReleaseParticles (:1534-1629), mcInletOutletBoundary::correct
// (mcBoundary/mcInletOutletBoundary/mcInletOutletBoundary.C:232-354).
#include <algorithm>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>

#include "cloud.hpp"
#include "evolve_kernel.cuh"

void fieldsFinalise(Cloud &c, double *partSum, double *partMax, int grid);   // fields.cu
void commAllreduceMoments(Cloud &c, float *ms);                                // comm.cu

// ---------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------
__global__ void k_iota(uint32_t *v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

__global__ void k_locate(DMesh M, int n, PState S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    S.tet[i] = locateTet(M, mk(S.px[i], S.py[i], S.pz[i]), S.cell[i]);
}

// fold per-CTA partial rows in index order (one thread per column)
// launched with (32, 8) threads: warp w folds rows w, w+8, ... of column threadIdx.x, then
// the eight partials are combined in warp order (fixed order -> deterministic)
#define FOLD_Y 32
__global__ void k_fold_rows(int nRows, int nCols, const double *rows, double *out, int isMax) {
    // blockDim = (32, FOLD_Y): thread (k, w) folds rows w, w + FOLD_Y, ... of column k, then row 0 of the block folds the
    // FOLD_Y partial results in order. One CTA: the inputs are a few hundred rows, the kernel is a chain of latencies
    // (93 dependent trips with 8 rows of threads on the 740 rows of the particle kernel; 24 with 32).
    __shared__ double part[FOLD_Y][33];
    const int k = threadIdx.x, w = threadIdx.y;
    double x = isMax ? -PDF_VGREAT : 0.0;
    if (k < nCols)
        for (int r = w; r < nRows; r += FOLD_Y) { const double y = rows[(size_t)r * nCols + k]; x = isMax ? fmax(x, y) : x + y; }
    part[w][k] = x;
    __syncthreads();
    if (w == 0 && k < nCols) {
        double t = part[0][k];
        for (int j = 1; j < FOLD_Y; ++j) t = isMax ? fmax(t, part[j][k]) : t + part[j][k];
        out[k] = t;
    }
}

// multi-block exclusive scan, phase 1: sum of each tile of SCAN_TILE elements
#define SCAN_TILE 4096
__global__ void k_scan_tile_sums(int n, const int *in, int *tileSums) {
    __shared__ int warpSums[32];
    const int base = blockIdx.x * SCAN_TILE;
    int s = 0;
    for (int j = threadIdx.x; j < SCAN_TILE; j += blockDim.x) { const int i = base + j; s += i < n ? in[i] : 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = threadIdx.x < (blockDim.x >> 5) ? warpSums[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) tileSums[blockIdx.x] = t;
    }
}
// phase 3: rescan every tile starting from its scanned tile offset (blockDim = 1024, 4 per thread)
__global__ void k_scan_tiles(int n, const int *in, const int *tileOffsets, int *out) {
    __shared__ int warpSums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    int v[4], t = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (i0 + k) < n ? in[i0 + k] : 0; t += v[k]; }
    int x = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warpSums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int s = warpSums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        warpSums[lane] = s;
    }
    __syncthreads();
    int run = tileOffsets[blockIdx.x] + (warp > 0 ? warpSums[warp - 1] : 0) + (x - t);
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (i0 + k < n) out[i0 + k] = run; run += v[k]; }
}

// exclusive scan of an int array in a single CTA (small n, or the tile sums of the multi-block scan)
__global__ void k_scan_single(int n, const int *in, int *out, int *total) {
    __shared__ int warpSums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warpSums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int s = lane < nw ? warpSums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            warpSums[lane] = s;
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? warpSums[warp - 1] : 0) + (x - v);
        if (i < n) out[i] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// ---------------------------------------------------------------------------
// segment bounds of the sorted keys and the per-cell moment reduction
// ---------------------------------------------------------------------------
// keys hold the cell's rank in the processing order; the bounds are stored per cell label
__global__ void k_cell_bounds(int n, const uint32_t *keys, int tetBits, const int *cellOfRank, int *cellStart, int *cellEnd,
                              StepScalars *sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    if (k == KEY_DEAD) { if (i == 0) sc->nSorted = 0; return; }
    const uint32_t r = k >> tetBits;
    if (i == 0 || (keys[i - 1] >> tetBits) != r) cellStart[cellOfRank[r]] = i;
    const uint32_t kn = (i + 1 < n) ? keys[i + 1] : KEY_DEAD;
    if (kn == KEY_DEAD || (kn >> tetBits) != r) cellEnd[cellOfRank[r]] = i + 1;
    if (kn == KEY_DEAD) sc->nSorted = i + 1;
}

// ---------------------------------------------------------------------------
// Counting sort by cell (used when the key is the cell label alone): histogram, exclusive scan,
// scatter — warp-aggregated integer atomics, so the *set* of entries of a cell is exact but
// their order depends on timing — and then every cell's segment is put into ascending order
// of the previous index by a bitonic network in registers. That is exactly the permutation a
// stable sort by cell produces (what cub::DeviceRadixSort::SortPairs gave in three 8-bit passes
// over 1 GB each), at one pass over the keys per kernel.
// ---------------------------------------------------------------------------
// (forms the key — the rank of the particle's new cell in the processing order — from the state the particle
// kernel has just written, and keeps it for the scatter)
// (the entry count is read on the device - no host round trip in the middle of a step; the trip count of the
// grid-stride loop is uniform per CTA, so every warp enters the match with all its lanes)
__global__ void k_cs_hist(const StepScalars *sc, const int *cell, const int *cellRank, uint32_t *keys, int *count) {
    const int n = sc->nEntries;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        uint32_t key = KEY_DEAD;
        if (i < n) {
            const int c = cell[i];
            if (c >= 0) key = (uint32_t)__ldg(cellRank + c);
            keys[i] = key;
        }
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (key != KEY_DEAD && (int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(count + key, __popc(m));
    }
}

__global__ void k_cs_scatter(const StepScalars *sc, const uint32_t *keys, const int *cellStart, int *cursor, uint32_t *perm) {
    const int n = sc->nEntries;
    const int lane = threadIdx.x & 31;
    for (int base0 = blockIdx.x * blockDim.x; base0 < n; base0 += gridDim.x * blockDim.x) {
        const int i = base0 + threadIdx.x;
        const uint32_t key = i < n ? keys[i] : KEY_DEAD;
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (key != KEY_DEAD && lane == leader) base = atomicAdd(cursor + key, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (key != KEY_DEAD) perm[cellStart[key] + base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    }
}

// bitonic sort of 32*K values held K per lane (element index = k*32 + lane), ascending
template <int K>
__device__ __forceinline__ void warpBitonic(uint32_t (&v)[K], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * K; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int ks = stride >> 5;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if ((k & ks) == 0) {
                        const bool asc = ((k << 5) & size) == 0;
                        const uint32_t a = v[k], b = v[k | ks];
                        const uint32_t lo = min(a, b), hi = max(a, b);
                        v[k] = asc ? lo : hi; v[k | ks] = asc ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint32_t p = __shfl_xor_sync(0xffffffffu, v[k], stride);
                    const bool asc = (((k << 5) | lane) & size) == 0;
                    const bool lower = (lane & stride) == 0;
                    v[k] = (lower == asc) ? min(v[k], p) : max(v[k], p);
                }
            }
        }
    }
}

template <int K>
__device__ __forceinline__ void sortSegment(uint32_t *seg, int n, int lane) {
    uint32_t v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; v[k] = j < n ? seg[j] : 0xffffffffu; }
    warpBitonic<K>(v, lane);
#pragma unroll
    for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; if (j < n) seg[j] = v[k]; }
}

// canonical (ascending) order of one cell's segment of the permutation, by one warp
__device__ __forceinline__ void orderSegment(uint32_t *seg, int n, int lane, uint32_t *scratchSeg) {
    if (n <= 1) return;
    if (n <= 32) sortSegment<1>(seg, n, lane);
    else if (n <= 64) sortSegment<2>(seg, n, lane);
    else if (n <= 128) sortSegment<4>(seg, n, lane);
    else if (n <= 256) sortSegment<8>(seg, n, lane);
    else {
        // very full cells: rank by counting through the scratch array (values are distinct)
        for (int j = lane; j < n; j += 32) {
            const uint32_t x = seg[j];
            int rk = 0;
            for (int q = 0; q < n; ++q) rk += seg[q] < x ? 1 : 0;
            scratchSeg[rk] = x;
        }
        __syncwarp();
        for (int j = lane; j < n; j += 32) seg[j] = scratchSeg[j];
        __syncwarp();
    }
}

// one warp per cell: segment end + canonical order of the segment
__global__ void k_cs_order(int nCells, const int *rankStart, const int *count, const int *cellOfRank, int *cellStart, int *cellEnd,
                           uint32_t *perm, uint32_t *scratch) {
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    for (int r = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); r < nCells; r += gridDim.x * warpsPerBlock) {
        const int s = rankStart[r], n = count[r];
        if (lane == 0) { const int c = cellOfRank[r]; cellStart[c] = s; cellEnd[c] = s + n; }
        orderSegment(perm + s, n, lane, scratch + s);
    }
}

// sort key of the radix path: (rank of the cell << tetBits) | local tet
__global__ void k_make_keys(DMesh M, int n, const int *cell, const int *tet, uint32_t *keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell[i];
    uint32_t key = KEY_DEAD;
    if (c >= 0) {
        key = (uint32_t)__ldg(M.cellRank + c);
        if (M.tetBits > 0) key = (key << M.tetBits) | (uint32_t)(tet[i] - __ldg(M.cellTetStart + c));
    }
    keys[i] = key;
}

__global__ void k_cs_finish(const int *total, StepScalars *sc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) sc->nSorted = *total;
}

template <int NACC, int G0>
__device__ __forceinline__ void reduceStoreGroups(const double (&a)[NACC], int lane, double *out) {
    if constexpr (G0 < NACC) {
        constexpr int LEFT = (NACC - G0 < 32) ? NACC - G0 : 32;
        constexpr int C = pow2Ceil(LEFT);
        constexpr int SHIFT = log2Floor(32 / C);
        double v[C];
#pragma unroll
        for (int k = 0; k < C; ++k) v[k] = (k < LEFT) ? a[G0 + (k < LEFT ? k : 0)] : 0.0;
        const double x = laneSums<C>(v, lane);
        const int k = lane >> SHIFT;
        if ((lane & ((1 << SHIFT) - 1)) == 0 && k < LEFT) out[G0 + k] = x;
        reduceStoreGroups<NACC, G0 + 32>(a, lane, out);
    }
}

// updateCloudPDF first half (mcParticleCloud.C:929-956): one warp per cell,
// lanes stride over the cell's segment of the sorted order, butterfly reduction
// in a fixed order -> deterministic sums without atomics.
// Block layout per cell: N, M, V, MU(3), MPhi(nPhi), MPhiPhi(nCov), MUU(6), MUPhi(3 nPhi); the last group is the
// velocity-scalar flux moment (BASELINE north_star "scalar fluxes"; the reference's moments stop at MUU).
// Compiled for 4 CTAs per SM (64 registers) up to one scalar, where the kernel is latency-bound; the wider
// instantiations hold 20 + 6 nPhi + nCov accumulators and would spill at that limit.
#define MOM_MIN_BLOCKS(NPHI) ((NPHI) <= 1 ? 4 : (NPHI) <= 3 ? 2 : 1)
// moments of cell c from its segment [s, e) of the sorted order, by one warp
template <int NPHI>
__device__ __forceinline__ void momentsOfSegment(const DMesh &M, const PState &S, const uint32_t *perm, int c, int s, int e, int lane,
                                                 double *block, int nMom) {
    constexpr int NCOV = NPHI * (NPHI + 1) / 2;
    constexpr int NACC = 11 + NPHI + NCOV + 3 * NPHI;   // all but the count
    double a[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) a[k] = 0.0;
    struct Rec { V3 U; double mpd, rho; double ph[NPHI > 0 ? NPHI : 1]; };
    auto fetch = [&](int j, Rec &R) {
        const int slot = (int)perm[j];
        R.U = mk(S.ux[slot], S.uy[slot], S.uz[slot]);
        double mpd = S.m[slot];
        if (M.axisym) mpd = massPerDepth(M, mk(S.px[slot], S.py[slot], S.pz[slot]), mpd);
        R.mpd = S.eta[slot] * mpd;
        R.rho = S.rho[slot];
#pragma unroll
        for (int k = 0; k < NPHI; ++k) R.ph[k] = S.phi[k][slot];
    };
    auto accumulate = [&](const Rec &R) {
        const double mpd = R.mpd;
        const V3 U = R.U;
        a[0] += mpd;
        a[1] += mpd / R.rho;
        a[2] += mpd * U.x; a[3] += mpd * U.y; a[4] += mpd * U.z;
#pragma unroll
        for (int k = 0; k < NPHI; ++k) a[5 + k] += mpd * R.ph[k];
        int pp = 0;
#pragma unroll
        for (int k = 0; k < NPHI; ++k)
#pragma unroll
            for (int l = k; l < NPHI; ++l, ++pp) a[5 + NPHI + pp] += mpd * R.ph[k] * R.ph[l];
        double *uu = a + 5 + NPHI + NCOV;
        uu[0] += mpd * (U.x * U.x); uu[1] += mpd * (U.x * U.y); uu[2] += mpd * (U.x * U.z);
        uu[3] += mpd * (U.y * U.y); uu[4] += mpd * (U.y * U.z); uu[5] += mpd * (U.z * U.z);
#pragma unroll
        for (int k = 0; k < NPHI; ++k) {
            const double mp = mpd * R.ph[k];
            uu[6 + 3 * k] += mp * U.x; uu[7 + 3 * k] += mp * U.y; uu[8 + 3 * k] += mp * U.z;
        }
    };
    // (two entries per lane and trip, both gathers requested before either is accumulated, measured slower:
    // 2.66 against 2.27 ms at 64 particles per cell - the registers it needs cost a CTA per SM)
    for (int j = s + lane; j < e; j += 32) {
        Rec R0;
        fetch(j, R0);
        accumulate(R0);
    }
    double *b = block + (size_t)c * nMom;
    if (lane == 0) b[0] = (double)(e - s);
    reduceStoreGroups<NACC, 0>(a, lane, b + 1);
}

template <int NPHI>
__global__ void __launch_bounds__(256, MOM_MIN_BLOCKS(NPHI)) k_moments(DMesh M, PState S, const uint32_t *perm, const int *cellStart, const int *cellEnd,
                          double *block, int nMom) {
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    for (int r = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); r < M.nCells; r += gridDim.x * warpsPerBlock) {
        const int c = M.cellOfRank[r];   // cells in processing order: neighbouring segments of the sorted arrays
        momentsOfSegment<NPHI>(M, S, perm, c, cellStart[c], cellEnd[c], lane, block, nMom);
    }
}

// k_cs_order and k_moments in one pass over the cells (counting-sort path without pair mixing): the warp that has put
// a cell's segment into its canonical order takes the cell's moments from it straight away - one launch, one read of
// the per-cell tables and the segment still in L1 instead of a second pass over the permutation
template <int NPHI>
__global__ void __launch_bounds__(256, MOM_MIN_BLOCKS(NPHI)) k_order_moments(DMesh M, PState S, const int *rankStart, const int *count, int *cellStart,
                                int *cellEnd, uint32_t *perm, uint32_t *scratch, double *block, int nMom) {
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    for (int r = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); r < M.nCells; r += gridDim.x * warpsPerBlock) {
        const int c = M.cellOfRank[r];
        const int s = rankStart[r], n = count[r];
        if (lane == 0) { cellStart[c] = s; cellEnd[c] = s + n; }
        orderSegment(perm + s, n, lane, scratch + s);
        __syncwarp();   // the segment as the other lanes have written it
        momentsOfSegment<NPHI>(M, S, perm, c, s, s + n, lane, block, nMom);
    }
}

// ---------------------------------------------------------------------------
// "modifiedCurl" pair mixing (new, not in the reference; definition in
// include/pdfb200.h, CPU restatement oracle/src/cloud.cpp Cloud::curlMixing).
// One warp per cell on the freshly sorted order. Per round the cell's particles
// are ranked by (Philox hash of the id, id) and consecutive ranks form disjoint
// pairs, so the pairs of a round are independent and one lane mixes one pair.
// Cells of up to 128 local particles are ranked through warp shuffles from
// registers; larger ones go through the two scratch arrays.
// ---------------------------------------------------------------------------
#define MIX_RMAX 8
#define MIX_WARPS 8
__global__ void __launch_bounds__(MIX_WARPS * 32) k_mix_curl(DMesh M, pdfb200_params P, PState S, const uint32_t *perm, const int *cellStart,
                                                              const int *cellEnd, const StepScalars *sc, RngKey key, uint32_t *hashScratch,
                                                              uint32_t *rankSlot, double *balancePartial) {
    __shared__ int slotByRank[MIX_WARPS][128];
    __shared__ double balSm[K1_MAXCONS * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // what the mixing adds to the step's balance of the conserved scalars, sum of massPerDepth x (phi_new - phi_old):
    // a pair conserves its eta*m-weighted mean, the balance (mcParticleCloud.C:1432-1444) weighs with m per depth
    double bal[K1_MAXCONS];
#pragma unroll
    for (int k = 0; k < K1_MAXCONS; ++k) bal[k] = 0.0;
    const uint32_t step = sc->step;
    const double deltaT = sc->deltaT;
    for (int c = blockIdx.x * MIX_WARPS + warp; c < M.nCells; c += gridDim.x * MIX_WARPS) {
        const int s = cellStart[c], e = cellEnd[c];
        const int n = e - s;
        if (n < 2) continue;
        double wSum = 0, wMax = 0, oeMax = 0;
        for (int j = s + lane; j < e; j += 32) {
            const int slot = (int)perm[j];
            const double eta = S.eta[slot], w = eta * S.m[slot];
            wSum += w; wMax = fmax(wMax, w); oeMax = fmax(oeMax, S.omega[slot] * eta);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wSum += __shfl_xor_sync(0xffffffffu, wSum, o);
            wMax = fmax(wMax, __shfl_xor_sync(0xffffffffu, wMax, o));
            oeMax = fmax(oeMax, __shfl_xor_sync(0xffffffffu, oeMax, o));
        }
        const double wBar = wSum / n;
        const double pc = 3.0 * P.Cmix * deltaT * oeMax * (wMax / wBar);
        const int R = min(MIX_RMAX, max(1, (int)ceil(pc)));
        auto mixPair = [&](int p, int q, int r) {
            double u, a;
            rngUniform2(key, S.id[p], step, RNG_MIX_PAIR, (uint32_t)r, u, a);
            const double etap = S.eta[p], etaq = S.eta[q];
            const double wp = etap * S.m[p], wq = etaq * S.m[q];
            const double Om = 0.5 * (S.omega[p] + S.omega[q]), et = 0.5 * (etap + etaq);
            const double Ppq = 3.0 * P.Cmix * Om * et * deltaT * ((wp + wq) / (2.0 * wBar)) / R;
            if (!(u < Ppq)) return;
            for (int mI = 0; mI < P.nMixed; ++mI) {
                double *col = S.phi[P.mixed[mI]];
                const double fp = col[p], fq = col[q];
                const double mean = (wp * fp + wq * fq) / (wp + wq);
                const double np_ = fp + a * (mean - fp), nq = fq + a * (mean - fq);
                col[p] = np_;
                col[q] = nq;
                for (int cs = 0; cs < P.nConserved; ++cs)
                    if (P.conserved[cs] == P.mixed[mI]) {
                        const double mpdP = massPerDepth(M, mk(S.px[p], S.py[p], S.pz[p]), S.m[p]);
                        const double mpdQ = massPerDepth(M, mk(S.px[q], S.py[q], S.pz[q]), S.m[q]);
                        const double d = mpdP * (np_ - fp) + mpdQ * (nq - fq);
#pragma unroll
                        for (int k = 0; k < K1_MAXCONS; ++k) if (k == cs) bal[k] += d;
                    }
            }
        };
        auto hashOf = [&](uint64_t id, int r) {
            uint32_t o[4];
            rngWords(key, id, step, RNG_MIX_HASH, (uint32_t)r, o);
            return o[0];
        };
        for (int r = 0; r < R; ++r) {
            if (n <= 128) {
                uint32_t hv[4];
                uint64_t idv[4];
                int slotv[4], rkv[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int j = s + lane + 32 * it;
                    const bool valid = j < e;
                    const int slot = valid ? (int)perm[j] : 0;
                    slotv[it] = slot;
                    idv[it] = valid ? S.id[slot] : ~0ull;
                    hv[it] = valid ? hashOf(idv[it], r) : 0xffffffffu;
                    rkv[it] = 0;
                }
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int nr = min(32, n - 32 * rr);   // warp-uniform
                    for (int l = 0; l < nr; ++l) {
                        const uint32_t hq = __shfl_sync(0xffffffffu, hv[rr], l);
                        const uint64_t iq = __shfl_sync(0xffffffffu, idv[rr], l);
#pragma unroll
                        for (int it = 0; it < 4; ++it) rkv[it] += (hq < hv[it] || (hq == hv[it] && iq < idv[it])) ? 1 : 0;
                    }
                }
#pragma unroll
                for (int it = 0; it < 4; ++it)
                    if (s + lane + 32 * it < e) slotByRank[warp][rkv[it]] = slotv[it];
                __syncwarp();
                for (int j = lane; 2 * j + 1 < n; j += 32) mixPair(slotByRank[warp][2 * j], slotByRank[warp][2 * j + 1], r);
                __syncwarp();
            } else {
                for (int j = s + lane; j < e; j += 32) hashScratch[j] = hashOf(S.id[perm[j]], r);
                __syncwarp();
                for (int j = s + lane; j < e; j += 32) {
                    const int slot = (int)perm[j];
                    const uint32_t h = hashScratch[j];
                    const uint64_t id = S.id[slot];
                    int rk = 0;
                    for (int q = s; q < e; ++q) {
                        const uint32_t hq = hashScratch[q];
                        if (hq < h || (hq == h && S.id[perm[q]] < id)) ++rk;
                    }
                    rankSlot[s + rk] = (uint32_t)slot;
                }
                __syncwarp();
                for (int j = lane; 2 * j + 1 < n; j += 32) mixPair((int)rankSlot[s + 2 * j], (int)rankSlot[s + 2 * j + 1], r);
                __syncwarp();
            }
        }
    }
    __syncthreads();
    blockReduce<K1_MAXCONS, false>(bal, balSm);
    if (threadIdx.x == 0)
        for (int k = 0; k < K1_MAXCONS; ++k) balancePartial[blockIdx.x * K1_MAXCONS + k] = bal[k];
}

// ---------------------------------------------------------------------------
// initReleaseParticles (mcParticleCloud.C:1534-1629): one thread per cell
// ---------------------------------------------------------------------------
__global__ void k_seed(DMesh M, pdfb200_params P, PState S, RngKey key, int rank, int nRanks, const double *rho_c,
                       const double *Ufv_c, const double *kfv_c, const double *Phi_c) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.nCells) return;
    const int Npc = P.particlesPerCell;
    const int nLoc = (Npc - rank + nRanks - 1) / nRanks;
    const double m = M.cellVolumes[c] * rho_c[c] / Npc;
    const V3 Updf = ld3(Ufv_c, c);
    const double urms = sqrt(2. / 3. * kfv_c[c]);
    const V3 ax = mk(M.axis[0], M.axis[1], M.axis[2]);
    double sumR = 0.;
    size_t o = (size_t)c * nLoc;
    for (int j = rank; j < Npc; j += nRanks, ++o) {
        const uint64_t idx = (uint64_t)((long long)c * Npc + j);
        const V3 pos = randomPointInCell(M, key, c, idx, RNG_SEED, 0);
        double g0, g1, g2, g3;
        rngNormal2(key, idx, 0, RNG_SEED, 1000000, g0, g1);
        rngNormal2(key, idx, 0, RNG_SEED, 1000001, g2, g3);
        const V3 U = mk(g0 * urms, g1 * urms, g2 * urms) + Updf;
        S.px[o] = pos.x; S.py[o] = pos.y; S.pz[o] = pos.z;
        S.ux[o] = U.x; S.uy[o] = U.y; S.uz[o] = U.z;
        S.m[o] = m; S.omega[o] = 0; S.rho[o] = 0; S.eta[o] = 1;
        for (int k = 0; k < P.nPhi; ++k) S.phi[k][o] = Phi_c[(size_t)k * M.nCells + c];
        S.cell[o] = c; S.tet[o] = locateTet(M, pos, c); S.id[o] = idx;
        if (M.axisym) sumR += mag3(pos - dot3(pos, ax) * ax);
    }
    if (M.axisym) {   // adjustAxiSymmetricMass (:2095-2117)
        const double alpha = nLoc / sumR;
        for (size_t q = (size_t)c * nLoc; q < (size_t)(c + 1) * nLoc; ++q) {
            const V3 pos = mk(S.px[q], S.py[q], S.pz[q]);
            S.m[q] *= alpha * mag3(pos - dot3(pos, ax) * ax);
        }
    }
}

// localTimeStepping/Omega/reaction correct() on freshly created particles
// (mcParticleCloud.C:1554-1561)
__global__ void k_init_models(DMesh M, pdfb200_params P, PState S, int n, const double *nodeMid, const double *cellAux,
                              int auxStride, const StepScalars *sc, FlameletTables FT) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 pos = mk(S.px[i], S.py[i], S.pz[i]);
    const int tet = S.tet[i], cell = S.cell[i];
    TetL T;
    loadTet(M, tet, T);
    double w0, w1, w2, w3;
    baryAt(T, pos - cellCentre(M, cell), w0, w1, w2, w3);
    double val[NODE_MID_STRIDE], pst[4];
    const int2 pts = tetPtsOf(M.tetD[tet]);
    interpMid(M, P, nodeMid, T.face, pts.x, pts.y, cell, w0, w1, w2, w3, val, pst);
    double eta = 1.0;
    if (P.localTimeStepping == PDFB200_LTS_CELL) eta = val[NM_ETA];
    else if (P.localTimeStepping == PDFB200_LTS_PARTICLE) {
        V3 Ut = mk(S.ux[i], S.uy[i], S.uz[i]);
        constrainDir(MESH_GEOD3(M), Ut);
        eta = fmin(P.CFL / stabilise(sc->deltaT * courantNo(M, cell, Ut), PDF_VSMALL), P.ltsUpperBound);
    }
    S.eta[i] = eta;
    S.m[i] = S.m[i] / eta;
    double Omega = S.omega[i];
    if (P.omegaModel != PDFB200_OMEGA_NONE) Omega = fmax(val[NM_OMEGA], PDF_SMALL);
    S.omega[i] = Omega;
    double phi[PDFB200_MAX_PHI];
#pragma unroll
    for (int k = 0; k < PDFB200_MAX_PHI; ++k) phi[k] = k < P.nPhi ? S.phi[k][i] : 0.0;
    double rho = S.rho[i];
    double CoUnused = 0;
    reactionCorrect(P, FT, Omega, val[NM_ZVAR], cellAux[(size_t)cell * auxStride + 1], eta * sc->deltaT, phi, rho, CoUnused);
    S.rho[i] = rho;
#pragma unroll
    for (int k = 0; k < PDFB200_MAX_PHI; ++k) if (k < P.nPhi) S.phi[k][i] = phi[k];
}

// ---------------------------------------------------------------------------
// boundaries before the move
// ---------------------------------------------------------------------------
// Un_ = n & U_patch on reflecting open patches (mcOpenBoundary.C:92-99)
__global__ void k_open_un(DMesh M, const double *Ufv_b, double *Un) {
    const int bf = blockIdx.x * blockDim.x + threadIdx.x;
    if (bf >= M.nBnd) return;
    const int info = M.bfInfo[bf], h = BF_HANDLER(info);
    if ((h == PDFB200_BC_OPEN || h == PDFB200_BC_INLET_OUTLET) && BF_REFLECTING(info)) {
        const double4 n = M.bfNormal[bf];
        Un[bf] = dot3(mk(n.x, n.y, n.z), ld3(Ufv_b, bf));
    }
}

// mcInletOutletBoundary ctor + getU()/getR() (.C:57-151): transforms, fan-triangle
// probabilities, and the inflow U / R in the face-normal frame, cached at first use
__global__ void k_inlet_cache(DMesh M, double k0, const double *Ufv_b, const double *Tau_b, double *fwdTrans,
                              double *probVec, double *inletU, double *inletR) {
    const int bf = blockIdx.x * blockDim.x + threadIdx.x;
    if (bf >= M.nBnd) return;
    if (BF_HANDLER(M.bfInfo[bf]) != PDFB200_BC_INLET_OUTLET) return;
    const int f = M.nInternalFaces + bf;
    const int s = M.faceStart[f], n = M.faceStart[f + 1] - s;
    const int s0 = M.faceStart[M.nInternalFaces];
    const V3 C = ld3(M.faceCentres, f);
    double area = 0;
    for (int i1 = 0, i2 = 1; i1 < n; ++i1, ++i2) {
        if (i2 == n) i2 = 0;
        const V3 v1 = ld3(M.points, M.facePoints[s + i1]) - C, v2 = ld3(M.points, M.facePoints[s + i2]) - C;
        area += mag3(cross3(v1, v2));
        probVec[s - s0 + i1] = area;
    }
    for (int i = 0; i < n; ++i) probVec[s - s0 + i] /= area;
    const double4 nn = M.bfNormal[bf];
    double T[9];
    rotationTensor(mk(-nn.x, -nn.y, -nn.z), mk(1.0, 0.0, 0.0), T);
    for (int k = 0; k < 9; ++k) fwdTrans[9 * (size_t)bf + k] = T[k];
    st3(inletU, bf, xform(T, ld3(Ufv_b, bf)));
    const double *tb = Tau_b + 6 * (size_t)bf;
    const double Sm[9] = {tb[0], tb[1], tb[2], tb[1], tb[3], tb[4], tb[2], tb[4], tb[5]};
    double TS[9], R[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double v = 0; for (int k = 0; k < 3; ++k) v += T[3 * i + k] * Sm[3 * k + j]; TS[3 * i + j] = v; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double v = 0; for (int k = 0; k < 3; ++k) v += TS[3 * i + k] * T[3 * j + k]; R[3 * i + j] = v; }
    double *r = inletR + 6 * (size_t)bf;
    r[0] = fmax(R[0], k0); r[1] = fmax(R[1], -PDF_GREAT); r[2] = fmax(R[2], -PDF_GREAT);
    r[3] = fmax(R[4], k0); r[4] = fmax(R[5], -PDF_GREAT); r[5] = fmax(R[8], k0);
}

// mcInletRandom (mcInletRandom.C:92-108, mcInletRandomI.H:28-39) + inversion
// generator (mcInversionInletRandom.C:63-119)
struct InletRnd {
    double UMean, b, b2, expmU2b2, UbSqrtPi, erfUb, denom, b22denom, Q, m;
    __device__ void updateCoeffs(double UMean_, double uRms) {
        UMean = UMean_;
        b = 0.70710678118654752440 / uRms;
        b2 = b * b;
        expmU2b2 = exp(-UMean * UMean * b2);
        UbSqrtPi = UMean * b * sqrt(3.14159265358979323846);
        erfUb = erf(UMean * b);
        denom = expmU2b2 + UbSqrtPi * (1. + erfUb);
        b22denom = 2. * b2 / denom;
        Q = 0.5 * (UMean + expmU2b2 / (b * sqrt(3.14159265358979323846)) + UMean * erfUb);
        m = 0.5 * (UMean * b + sqrt(2.0 + UMean * UMean * b2)) / b;
    }
    __device__ double PDF(double x) const { const double xp = x - UMean; return b22denom * exp(-b2 * xp * xp) * x; }
    __device__ double CDF(double x) const { x -= UMean; return (expmU2b2 - exp(-b2 * x * x) + UbSqrtPi * (erfUb + erf(b * x))) / denom; }
    __device__ double value(double U01) const {
        const double d = ((U01 - CDF(m)) >= 0 ? 1.0 : -1.0) * PDF_SMALL;
        double x0 = m + d;
        int nIter = 0;
        do {
            const double fval = CDF(x0) - U01, fder = PDF(x0);
            if (fabs(fder) < PDF_VSMALL) return x0;
            const double x = x0 - fval / fder;
            if (fabs(x - x0) < 1e-8) return x;
            x0 = x;
            ++nIter;
        } while (nIter < 50);
        return x0;
    }
    // "sampling" (mcSamplingInletRandom.C:37-102): x = UMean + uRms N(0,1) is accepted with a probability proportional
    // to x. The reference scales with the largest |x| of a batch of 1e6 deviates (and a factor 0.1); here the scale is
    // the largest value the generator can produce (its normals end at 6.66 sigma), so no state is needed and every
    // (face, step, draw) has its own stream. Same target density x exp(-(x-U)^2 b^2)/Q.
    __device__ double sample(RngKey key, uint32_t bf, uint32_t step, uint32_t kDraw, double uRms) const {
        const double xMax = fmax(UMean + 6.7 * uRms, PDF_SMALL);
        const uint64_t entity = (uint64_t)bf | ((uint64_t)kDraw << 32);   // one stream per (face, draw); sub = attempt
        for (uint32_t attempt = 0; attempt < (1u << 20); ++attempt) {
            uint32_t o[4];
            rngWords(key, entity, step, RNG_INJECT_SAMPLE, attempt, o);
            float n0, n1;
            normalPairF32(o[0], o[1], n0, n1);
            const double u = (double)((((unsigned long long)o[3] << 32) | o[2]) >> 11) * 0x1.0p-53;
            const double x = UMean + uRms * (double)n0;
            if (u * xMax < x) return x;
        }
        return m;   // (not reached: the acceptance probability is above 5 %)
    }
};

// mcInletOutletBoundary::correct(false) (.C:232-354): one WARP per boundary face. The
// reference's loop `while (mGen < mIn)` is sequential only in the accumulated mass: the
// candidates themselves (inflow velocity by Newton inversion, point on the face, tet,
// local time step) are keyed by (face, k) and independent, so the 32 lanes generate
// candidates k = kBase + lane in parallel and then replay the accept/stop logic in k
// order through shuffles, with the same floating-point sequence as the serial loop.
// pass 0 counts; pass 1 regenerates the same particles (counter-based RNG) and writes
// them at the offsets of an exclusive scan, so slots and ids are deterministic.
__global__ void k_inject(DMesh M, pdfb200_params P, int pass, PState S, RngKey key, StepScalars *sc, const double *Un,
                         const double *rho_b, const double *Phi_b, const double *inletU, const double *inletR,
                         const double *fwdTrans, const double *probVec, const double *nodeMid, int *count,
                         const int *offset, int *injPatchTail, double *injSums, int nRanks, int rank, long long capacity) {
    const int lane = threadIdx.x & 31;
    const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int bf = warpGlobal; bf < M.nBnd; bf += nWarps) {
        const int info = M.bfInfo[bf];
        if (pass == 0 && lane == 0) count[bf] = 0;
        if (pass == 1 && lane < K1_NC1) injSums[(size_t)bf * K1_NC1 + lane] = 0;
        if (BF_HANDLER(info) != PDFB200_BC_INLET_OUTLET) continue;
        if ((bf % nRanks) != rank) continue;        // inflow faces are dealt round-robin to the ranks
        if (Un[bf] > 0) continue;                   // outlet face
        const int nGen = pass == 1 ? count[bf] : 0;
        if (pass == 1 && nGen == 0) continue;
        const int f = M.nInternalFaces + bf, cellI = M.owner[f];
        const double deltaT = sc->deltaT;
        const uint32_t step = sc->step;
        InletRnd inrnd;
        const V3 Uloc = ld3(inletU, bf);
        const double *R = inletR + 6 * (size_t)bf;
        inrnd.updateCoeffs(Uloc.x, sqrt(R[0]));
        const double mp = rho_b[bf] * M.cellVolumes[cellI] / P.particlesPerCell;
        const double mIn = rho_b[bf] * inrnd.Q * M.magSf[f] * deltaT;
        const double *T = fwdTrans + 9 * (size_t)bf;
        const int s = M.faceStart[f], n = M.faceStart[f + 1] - s, s0 = M.faceStart[M.nInternalFaces];
        const double *wv = probVec + (s - s0);
        const V3 C = ld3(M.faceCentres, f), Sf = ld3(M.faceAreas, f);
        const V3 ax = mk(M.axis[0], M.axis[1], M.axis[2]);
        const double sqRyy = sqrt(R[3]), sqRzz = sqrt(R[5]);
        double sumR = 0.;
        const int loops = (pass == 1 && M.axisym) ? 2 : 1;
        for (int loop = 0; loop < loops; ++loop) {
            const bool writing = pass == 1 && loop == loops - 1;
            double mGen = 0;
            int made = 0;
            double sums[K1_NC1];
#pragma unroll
            for (int k = 0; k < K1_NC1; ++k) sums[k] = 0;
            bool done = false;
            for (uint32_t kBase = 0; !done; kBase += 32) {
                const uint32_t kDraw = kBase + lane;
                double uVal, uTri, n1, n2, a1, a2, uAcc, dummy;
                rngUniform2(key, (uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 0, uVal, uTri);
                rngNormal2(key, (uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 1, n1, n2);
                rngUniform2(key, (uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 2, a1, a2);
                rngUniform2(key, (uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 3, uAcc, dummy);
                // randomVelocity (.C:215-229): U_global = revTrans & Up = fwdTrans^T & Up
                const double uIn = P.inletRandom == PDFB200_INLET_RANDOM_SAMPLING ? inrnd.sample(key, (uint32_t)bf, step, kDraw, sqrt(R[0])) : inrnd.value(uVal);
                const V3 UpL = mk(uIn, Uloc.y + sqRyy * n1, Uloc.z + sqRzz * n2);
                const V3 Up = mk(T[0] * UpL.x + T[3] * UpL.y + T[6] * UpL.z, T[1] * UpL.x + T[4] * UpL.y + T[7] * UpL.z,
                                 T[2] * UpL.x + T[5] * UpL.y + T[8] * UpL.z);
                // randomPoint (.C:179-212)
                int idx1 = 0;
                while (idx1 < n && wv[idx1] < uTri) ++idx1;   // lower_bound
                int idx2 = idx1 + 1;
                if (idx2 == n) idx2 = 0;
                if (idx1 == n) idx1 = n - 1;                  // uTri beyond the last entry (== 1 up to round-off)
                if (a1 + a2 > 1.0) { a1 = 1.0 - a1; a2 = 1.0 - a2; }
                V3 pos = ((a1 * ld3(M.points, M.facePoints[s + idx1]) + a2 * ld3(M.points, M.facePoints[s + idx2])) + (1.0 - a1 - a2) * C) - PDF_SMALL * Sf;
                if (M.nGeoD <= 2) constrainDir(MESH_GEOD3(M), pos);
                // local time stepping of the new particle
                const int tet = locateTet(M, pos, cellI);
                double eta = 1.0;
                if (P.localTimeStepping == PDFB200_LTS_CELL) {
                    TetL TL;
                    loadTet(M, tet, TL);
                    double w0, w1, w2, w3, val[NODE_MID_STRIDE], pst[4];
                    baryAt(TL, pos - cellCentre(M, cellI), w0, w1, w2, w3);
                    const int2 pts = tetPtsOf(M.tetD[tet]);
                    interpMid(M, P, nodeMid, TL.face, pts.x, pts.y, cellI, w0, w1, w2, w3, val, pst);
                    eta = val[NM_ETA];
                } else if (P.localTimeStepping == PDFB200_LTS_PARTICLE) {
                    V3 Ut = Up;
                    constrainDir(MESH_GEOD3(M), Ut);
                    eta = fmin(P.CFL / stabilise(deltaT * courantNo(M, cellI, Ut), PDF_VSMALL), P.ltsUpperBound);
                }
                const double m = mp / eta;
                const double rAx = M.axisym ? mag3(pos - dot3(pos, ax) * ax) : 0.0;
                // replay of the serial accept/stop logic in k order (all lanes compute the same scalars)
                int nAcc = 0;
                for (int j = 0; j < 32; ++j) {
                    const double mj = __shfl_sync(0xffffffffu, m, j), uj = __shfl_sync(0xffffffffu, uAcc, j);
                    if (!(mGen < mIn)) { done = true; break; }
                    if (mGen + mj > mIn) {
                        if (uj > (mIn - mGen) / mj) { done = true; break; }
                    }
                    if (pass == 1 && loop == 0 && M.axisym) sumR += __shfl_sync(0xffffffffu, rAx, j);
                    mGen += mj;
                    ++nAcc;
                }
                double mm = m, mpd = 0.0;
                const bool mine = lane < nAcc;
                if (writing) {
                    if (M.axisym) mm *= (nGen / sumR) * rAx;
                    mpd = massPerDepth(M, pos, mm);
                    if (mine) {
                        const long long tailIdx = (long long)sc->nInjStart + offset[bf] + made + lane;
                        const long long slot = (long long)sc->nSlots + tailIdx;
                        if (slot < capacity) {
                            S.px[slot] = pos.x; S.py[slot] = pos.y; S.pz[slot] = pos.z;
                            S.ux[slot] = Up.x; S.uy[slot] = Up.y; S.uz[slot] = Up.z;
                            S.m[slot] = mm; S.omega[slot] = 0; S.rho[slot] = 0; S.eta[slot] = eta;
                            for (int k = 0; k < P.nPhi; ++k) S.phi[k][slot] = Phi_b[(size_t)k * M.nBnd + bf];
                            S.cell[slot] = cellI; S.tet[slot] = tet;
                            S.id[slot] = sc->nextId + (uint64_t)(offset[bf] + made + lane) * (uint64_t)nRanks;
                            injPatchTail[tailIdx] = BF_PATCH(info);
                        }
                    }
                    // mass-in sums in k order
                    for (int j = 0; j < nAcc; ++j) {
                        const double mj = __shfl_sync(0xffffffffu, mpd, j);
                        sums[0] += mj;
                        for (int cs = 0; cs < P.nConserved; ++cs) sums[1 + cs] += mj * Phi_b[(size_t)P.conserved[cs] * M.nBnd + bf];
                    }
                }
                made += nAcc;
                if (nAcc < 32) done = true;
            }
            if (pass == 0 && lane == 0) count[bf] = made;
            if (writing && lane == 0)
                for (int k = 0; k < K1_NC1; ++k) injSums[(size_t)bf * K1_NC1 + k] = sums[k];
        }
    }
}


__global__ void k_inject_finish(int nBnd, const int *total, StepScalars *sc, long long capacity, int nRanks) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int t = *total;
    if ((long long)sc->nSlots + sc->nTail + t > capacity) { sc->overflow = 1; t = max(0, (int)(capacity - sc->nSlots - sc->nTail)); }
    sc->nTail += t;
    // ids are 32 bits wide: running out of them must stop the run, not hand two particles one random stream
    sc->nextId += (uint64_t)t * (uint64_t)nRanks;
}

// ---------------------------------------------------------------------------
// particle-number control (mcParticleCloud.C:1014-1221)
// ---------------------------------------------------------------------------
// The weight the clone selection ranks by: eta*m rounded to single precision. The reference sorts by the
// double product and leaves the order of equal weights to std::sort; freshly seeded or injected particles of a
// cell carry m = rho V / (Npc eta), i.e. products that agree to the last bit or two, so that order would hang on
// the round-off of the moment sums. Weights that agree to 2^-24 count as equal here and the id decides
// (same rule in oracle/src/cloud.cpp).
__device__ __forceinline__ double cloneWeight(double eta, double m) { return (double)__double2float_rn(eta * m); }
__global__ void k_nc_classify(DMesh M, pdfb200_params P, const double *PaNIC, const int *cellStart, const int *cellEnd,
                              int *flag, int *nClone, int nRanks, int *list, int *nList) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.nCells) return;
    const int Npc = P.particlesPerCell;
    const int np = (int)round(PaNIC[c]);
    int fl = 0, n = 0;
    if (np < 1) fl = -1;
    else if (np < round(Npc * P.cloneAt)) fl = 1;
    else if (np > round(Npc * P.eliminateAt)) fl = 2;
    if (fl == 1) {
        n = min(np, Npc - np);
        const int nLocal = cellEnd[c] - cellStart[c];
        // particle-sharded runs (no counterpart in the reference): every rank steers its own share of the cell towards
        // 1/nRanks of the population the cell has after cloning. (Shares proportional to the local counts have no
        // restoring force: the ranks' holdings drift apart like competing species - 6 % after 400 steps on 4 ranks.)
        if (nRanks > 1) n = (int)fmin((double)nLocal, fmax(0., floor((double)(np + n) / nRanks - nLocal + 0.5)));
        n = min(n, nLocal);
    }
    flag[c] = fl; nClone[c] = n;
    // cells that need work go on a list (its order does not matter: cells are independent)
    if (fl > 0) list[atomicAdd(nList, 1)] = c;
}

// one warp per flagged cell
__global__ void __launch_bounds__(256, 4) k_nc_apply(DMesh M, pdfb200_params P, PState S, const uint32_t *perm, const int *cellStart,
                           const int *cellEnd, const int *flag, const int *nClone, const int *cloneOffset,
                           double *PaNIC, StepScalars *sc, RngKey key, int rank, int nRanks, long long capacity,
                           double *ncDelta, int *ncCounts /* [2]: cloned, eliminated */, uint32_t *cloneSrc, const int *list,
                           const int *nListPtr) {
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const uint32_t step = sc->step;
    const int nList = *nListPtr;
    for (int li = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); li < nList; li += gridDim.x * warpsPerBlock) {
        const int c = list[li];
        const int fl = flag[c];
        if (fl <= 0) continue;
        const int s = cellStart[c], e = cellEnd[c];
        double dsum[K1_NC1];
#pragma unroll
        for (int k = 0; k < K1_NC1; ++k) dsum[k] = 0;
        const int cnt = e - s;
        // cells of up to 128 local particles (4 per lane) are handled from registers; larger ones
        // re-read global memory
        const bool inRegs = cnt <= 128;
        if (fl == 1) {
            // cloneParticles (:1107-1139): the n heaviest by eta*m (see cloneWeight; ties by id) are split
            const int n = nClone[c];
            // the clones themselves are created by k_nc_clone, one thread per task
            auto makeClone = [&](int slot, int rk, uint64_t) { cloneSrc[cloneOffset[c] + rk] = (uint32_t)slot; };
            if (inRegs) {
                double keyv[4]; uint64_t idv[4]; int slotv[4], rkv[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int j = s + lane + 32 * it;
                    const bool valid = j < e;
                    const int slot = valid ? (int)perm[j] : 0;
                    slotv[it] = slot;
                    keyv[it] = valid ? cloneWeight(S.eta[slot], S.m[slot]) : -PDF_VGREAT;
                    idv[it] = valid ? S.id[slot] : ~0ull;
                    rkv[it] = 0;
                }
                // rank every entry against all others through warp shuffles
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int nr = min(32, cnt - 32 * r);   // warp-uniform
                    for (int l = 0; l < nr; ++l) {
                        const double kq = __shfl_sync(0xffffffffu, keyv[r], l);
                        const uint64_t iq = __shfl_sync(0xffffffffu, idv[r], l);
#pragma unroll
                        for (int it = 0; it < 4; ++it) rkv[it] += (kq > keyv[it] || (kq == keyv[it] && iq < idv[it])) ? 1 : 0;
                    }
                }
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    if (s + lane + 32 * it < e && rkv[it] < n) {
                        makeClone(slotv[it], rkv[it], idv[it]);
                        S.m[slotv[it]] = S.m[slotv[it]] / 2.0;   // ranks came from registers: safe to halve now
                    }
                }
            } else {
                unsigned long long split[4] = {0ull, 0ull, 0ull, 0ull};   // this lane's entries that get halved (first 256)
                int it = 0;
                for (int j = s + lane; j < e; j += 32, ++it) {
                    const int slot = (int)perm[j];
                    const double key_i = cloneWeight(S.eta[slot], S.m[slot]);
                    const uint64_t id_i = S.id[slot];
                    int rk = 0;
                    for (int q = s; q < e; ++q) {
                        const int sq = (int)perm[q];
                        const double kq = cloneWeight(S.eta[sq], S.m[sq]);
                        if (kq > key_i || (kq == key_i && S.id[sq] < id_i)) ++rk;
                    }
                    if (rk < n && it < 256) {
                        split[it >> 6] |= 1ull << (it & 63);
                        makeClone(slot, rk, id_i);
                    }
                }
                __syncwarp();
                // halve the sources only after every lane has ranked against the old masses
                it = 0;
                for (int j = s + lane; j < e; j += 32, ++it) {
                    if (it < 256 && ((split[it >> 6] >> (it & 63)) & 1ull)) {
                        const int slot = (int)perm[j];
                        S.m[slot] = S.m[slot] / 2.0;
                    }
                }
            }
            if (lane == 0) { PaNIC[c] += n; atomicAdd(ncCounts, n); }
        } else {
            // eliminateParticles (:1143-1221)
            const int ncur = (int)round(PaNIC[c]);
            const int nx = ncur - P.particlesPerCell;
            double Psel = (2. * nx) / ncur;
            // particle-sharded runs: the rank removes what it holds above its share Npc / nRanks of the cell
            if (nRanks > 1) Psel = fmin(1., 2. * fmax(0., cnt - (double)P.particlesPerCell / nRanks) / max(cnt, 1));
            double mA = 0., mB = 0.;
            // selection state of this lane's first four entries (all of them when inRegs)
            int slotv[4]; double u2v[4]; bool selv[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int j = s + lane + 32 * it;
                selv[it] = false; slotv[it] = 0; u2v[it] = 0;
                if (j < e) {
                    const int slot = (int)perm[j];
                    double u1, u2;
                    rngUniform2(key, S.id[slot], step, RNG_NC_PARTICLE, 0, u1, u2);
                    slotv[it] = slot; u2v[it] = u2; selv[it] = u1 < Psel;
                    if (selv[it]) {
                        const double meta = S.eta[slot] * S.m[slot];
                        if (u2 < 0.5) mA += meta; else mB += meta;
                    }
                }
            }
            for (int j = s + lane + 128; j < e; j += 32) {
                const int slot = (int)perm[j];
                double u1, u2;
                rngUniform2(key, S.id[slot], step, RNG_NC_PARTICLE, 0, u1, u2);
                if (u1 < Psel) {
                    const double meta = S.eta[slot] * S.m[slot];
                    if (u2 < 0.5) mA += meta; else mB += meta;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { mA += __shfl_xor_sync(0xffffffffu, mA, o); mB += __shfl_xor_sync(0xffffffffu, mB, o); }
            if (mA > 0 || mB > 0) {
                const double Pd = mB / (mA + mB);
                double u, dummy;
                rngUniform2(key, (uint32_t)c, step, RNG_NC_CELL, (uint32_t)rank, u, dummy);
                const bool delA = u < Pd;
                const double sfac = delA ? 1. / Pd : (mA + mB) / mA;
                int nKilled = 0;
                auto settle = [&](int slot, bool inA) {
                    const V3 pos = mk(S.px[slot], S.py[slot], S.pz[slot]);
                    const double mOld = S.m[slot];
                    double d;
                    if (inA == delA) { S.cell[slot] = -1; ++nKilled; d = -massPerDepth(M, pos, mOld); }
                    else { const double mNew = mOld * sfac; S.m[slot] = mNew; d = massPerDepth(M, pos, mNew) - massPerDepth(M, pos, mOld); }
                    dsum[0] += d;
                    for (int cs = 0; cs < P.nConserved; ++cs) dsum[1 + cs] += d * S.phi[P.conserved[cs]][slot];
                };
#pragma unroll
                for (int it = 0; it < 4; ++it)
                    if (selv[it]) settle(slotv[it], u2v[it] < 0.5);
                for (int j = s + lane + 128; j < e; j += 32) {
                    const int slot = (int)perm[j];
                    double u1, u2;
                    rngUniform2(key, S.id[slot], step, RNG_NC_PARTICLE, 0, u1, u2);
                    if (u1 < Psel) settle(slot, u2 < 0.5);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) nKilled += __shfl_xor_sync(0xffffffffu, nKilled, o);
                if (lane == 0) { PaNIC[c] -= nKilled; atomicAdd(ncCounts + 1, nKilled); }
            }
        }
#pragma unroll
        for (int k = 0; k < K1_NC1; ++k) {
            double x = dsum[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) ncDelta[(size_t)c * K1_NC1 + k] = x;
        }
    }
}

// cloneParticles, second half (:1117-1137): one thread per clone task (source slot, already
// halved) creates the copy at a random point of the cell; task t lands in slot nSlots + t
__global__ void k_nc_clone(DMesh M, pdfb200_params P, PState S, const uint32_t *cloneSrc, const int *nTasks, StepScalars *sc,
                           RngKey key, int nRanks, long long capacity, double *cloneDelta) {
    const int n = *nTasks;
    const uint32_t step = sc->step;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const long long dst = (long long)sc->nSlots + t;
        if (dst >= capacity) { sc->overflow = 1; continue; }
        const int slot = (int)cloneSrc[t];
        const int c = S.cell[slot];
        const V3 posOld = mk(S.px[slot], S.py[slot], S.pz[slot]);
        const V3 pos = randomPointInCell(M, key, c, S.id[slot], RNG_CLONE_POS, step);
        const double mHalf = S.m[slot], mOld = 2.0 * mHalf;   // the source was halved by k_nc_apply (exact)
        S.px[dst] = pos.x; S.py[dst] = pos.y; S.pz[dst] = pos.z;
        S.ux[dst] = S.ux[slot]; S.uy[dst] = S.uy[slot]; S.uz[dst] = S.uz[slot];
        S.m[dst] = mHalf; S.omega[dst] = S.omega[slot]; S.rho[dst] = S.rho[slot]; S.eta[dst] = S.eta[slot];
        for (int k = 0; k < P.nPhi; ++k) S.phi[k][dst] = S.phi[k][slot];
        S.cell[dst] = c; S.tet[dst] = locateTet(M, pos, c);
        S.id[dst] = sc->nextId + (uint64_t)t * (uint64_t)nRanks;
        if (M.axisym) {
            const double d = massPerDepth(M, pos, mHalf) + massPerDepth(M, posOld, mHalf) - massPerDepth(M, posOld, mOld);
            cloneDelta[(size_t)t * K1_NC1] = d;
            for (int cs = 0; cs < K1_MAXCONS; ++cs) cloneDelta[(size_t)t * K1_NC1 + 1 + cs] = cs < P.nConserved ? d * S.phi[P.conserved[cs]][slot] : 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// end of step: lost-mass shares, report scalars, next deltaT
// ---------------------------------------------------------------------------
struct DevReport {
    double v[64];
};
// layout of DevReport.v
#define RP_K1SUM 0                       // ACC_NSUM entries
#define RP_K1MAX (RP_K1SUM + ACC_NSUM)   // ACC_NMAX entries
#define RP_CELLSUM (RP_K1MAX + ACC_NMAX) // 3: sum|rhoErr|, totalMass, totalParticleMass
#define RP_CELLMAX (RP_CELLSUM + 3)      // 2: -rhoErrMin, rhoErrMax
#define RP_NCDELTA (RP_CELLMAX + 2)      // K1_NC1
#define RP_INJ (RP_NCDELTA + K1_NC1)     // K1_NC1
#define RP_LOSTSUM (RP_INJ + K1_NC1)     // K1_NC1: lost mass re-spread, per depth, x {1, conserved scalars}
#define RP_CLONEDELTA (RP_LOSTSUM + K1_NC1)   // K1_NC1: axisymmetric mass-per-depth change of the clones
#define RP_MIXDELTA (RP_CLONEDELTA + K1_NC1)   // K1_MAXCONS: change of the conserved-scalar balance by pair mixing
#define RP_STEP (RP_MIXDELTA + K1_MAXCONS)     // 8: what the host used to read from StepScalars in the middle of the step
#define RP_DT_USED (RP_STEP + 0)
#define RP_DT_NEXT (RP_STEP + 1)
#define RP_N_ENTRIES (RP_STEP + 2)
#define RP_N_INJECTED (RP_STEP + 3)
#define RP_N_SORTED (RP_STEP + 4)              // alive after the sort
#define RP_N_CLONED (RP_STEP + 5)
#define RP_N_ELIMINATED (RP_STEP + 6)
#define RP_OVERFLOW (RP_STEP + 7)
#define RP_END (RP_STEP + 8)
static_assert(RP_END <= 64, "DevReport too small");

// Lost particles (mcParticleCloud.C:1257-1260,1346-1351, notifyLostParticle :2088-2092): the particle
// kernel appends (entry index, cell, mass) records to a list in whatever order its atomic counter
// hands out. This kernel (one CTA) puts the list into (cell, entry index) order by counting ranks and
// then sums every cell's records front to back, so lostMass does not depend on that order and no
// floating-point atomic is needed. The list is short (a healthy run loses nothing).
__global__ void k_lost_reduce(const LostRec *list, LostRec *sorted, const unsigned *count, unsigned cap, double *lostMass, StepScalars *sc) {
    unsigned n = *count;
    if (n == 0) return;
    if (n > cap) { n = cap; if (threadIdx.x == 0) sc->overflow = 2; }
    for (unsigned r = threadIdx.x; r < n; r += blockDim.x) {
        const LostRec me = list[r];
        const unsigned long long key = ((unsigned long long)(unsigned)me.cell << 32) | me.idx;
        unsigned rank = 0;
        for (unsigned q = 0; q < n; ++q) {
            const LostRec o = list[q];
            rank += ((((unsigned long long)(unsigned)o.cell << 32) | o.idx) < key) ? 1u : 0u;
        }
        sorted[rank] = me;
    }
    __syncthreads();
    for (unsigned r = threadIdx.x; r < n; r += blockDim.x) {
        const int c = sorted[r].cell;
        if (r > 0 && sorted[r - 1].cell == c) continue;   // not the head of its cell's run
        double sum = 0;
        for (unsigned q = r; q < n && sorted[q].cell == c; ++q) sum += sorted[q].m;
        lostMass[c] = sum;
    }
}

// lostMass_/max(PaNIC,1) (mcParticleCloud.C:1363-1369).
// The mass a rank lost is shared among the particles this rank holds in the cell after number
// control: nLocal = (segment length) + (PaNIC - N), where N is the all-reduced count of the moment
// block and PaNIC has since been updated by this rank's clones and eliminations. With one rank that
// is PaNIC itself (the reference's expression); with several it conserves the mass rank by rank.
__global__ void k_lost_share(int nCells, const double *lostMass, const double *PaNIC, const int *cellStart,
                             const int *cellEnd, const double *block, int nMom, double *lostShare) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x) {
        const double nLocal = (double)(cellEnd[c] - cellStart[c]) + (PaNIC[c] - block[(size_t)c * nMom]);
        lostShare[c] = lostMass[c] / fmax(nLocal, 1.);
    }
}

// What the re-spread adds to the step's mass balance (mcParticleCloud.C:1363-1369 runs before the sums of
// :1432-1444, so the reference's deltaMassInst sees the masses with their shares): the sum over the particles
// alive after number control of massPerDepth(share) x {1, conserved scalars}. The shares themselves are added to
// the masses by the next step's particle kernel. Nothing was lost in almost every step: the kernel then only
// writes zeros. Per-CTA partial rows, folded by k_fold_rows.
__global__ void k_lost_balance(DMesh M, pdfb200_params P, PState S, const uint32_t *perm, const StepScalars *sc, const unsigned *lostCount,
                               const double *lostShare, double *partial) {
    __shared__ double sm[K1_NC1 * 32];
    double v[K1_NC1];
#pragma unroll
    for (int k = 0; k < K1_NC1; ++k) v[k] = 0;
    if (*lostCount != 0) {
        const int nSorted = sc->nSorted, n = nSorted + sc->nTail, nSlots = sc->nSlots;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int slot = i < nSorted ? (int)perm[i] : nSlots + (i - nSorted);
            const int c = S.cell[slot];
            if (c < 0) continue;
            const double sh = lostShare[c];
            if (sh == 0.0) continue;
            const double mpd = massPerDepth(M, mk(S.px[slot], S.py[slot], S.pz[slot]), sh);
            v[0] += mpd;
            for (int cs = 0; cs < P.nConserved; ++cs) v[1 + cs] += mpd * S.phi[P.conserved[cs]][slot];
        }
    }
    blockReduce<K1_NC1, false>(v, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < K1_NC1; ++k) partial[blockIdx.x * K1_NC1 + k] = v[k];
}

// column sums of an [n][K1_NC1] array: per-CTA partial rows (folded by k_fold_rows)
__global__ void k_colsum_nc1(int n, const double *in, double *partial, const int *nDev = nullptr) {
    __shared__ double sm[K1_NC1 * 32];
    if (nDev) n = *nDev;
    double v[K1_NC1];
#pragma unroll
    for (int k = 0; k < K1_NC1; ++k) v[k] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < K1_NC1; ++k) v[k] += in[(size_t)i * K1_NC1 + k];
    blockReduce<K1_NC1, false>(v, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < K1_NC1; ++k) partial[blockIdx.x * K1_NC1 + k] = v[k];
}

// sums of the per-chunk mass sums of the particle kernel ([chunk][4 * K1_NC1]) in a fixed order: CTA b takes a fixed
// contiguous range of chunks, thread (sub, col) every 16th chunk of it, then the 16 subs are added in order
__global__ void k_chunk_fold(const StepScalars *sc, int chunk, const double *chunkSums, double *partial) {
    constexpr int NC = 4 * K1_NC1;
    __shared__ double sm[256];
    const int nChunks = (sc->nEntries + chunk - 1) / chunk;
    const int per = (nChunks + gridDim.x - 1) / gridDim.x;
    const int lo = min(nChunks, (int)blockIdx.x * per), hi = min(nChunks, lo + per);
    const int col = threadIdx.x % NC, sub = threadIdx.x / NC, nSub = blockDim.x / NC;
    double x = 0.0;
    for (int c = lo + sub; c < hi; c += nSub) x += chunkSums[(size_t)c * NC + col];
    sm[threadIdx.x] = x;
    __syncthreads();
    if (sub == 0) {
        double t = sm[col];
        for (int j = 1; j < nSub; ++j) t += sm[j * NC + col];
        partial[(size_t)blockIdx.x * NC + col] = t;
    }
}

__global__ void k_set_slots(StepScalars *sc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) sc->nSlots = sc->nEntries;
}

// after the injection: the entry count of the step, kept on the device for every later kernel; also empties the list
// of lost particles and notes what the report needs from the state at this point
__global__ void k_step_begin(StepScalars *sc, DevReport *rp, unsigned *lostCount, int *workCtr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    sc->nEntries = sc->nSorted + sc->nTail;
    *lostCount = 0u;
    *workCtr = 0;
    rp->v[RP_DT_USED] = sc->deltaT;
    rp->v[RP_N_ENTRIES] = (double)sc->nEntries;
    rp->v[RP_N_INJECTED] = (double)(sc->nTail - sc->nInjStart);
}

__global__ void k_step_end(StepScalars *sc, DevReport *rp, pdfb200_params P, int nTailClones, const int *ncCounts,
                           int nRanksForId) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double CoMax = rp->v[RP_K1MAX + ACC_COMAX];
    if (CoMax > 0) sc->deltaT = P.CFL / CoMax;
    sc->step += 1;
    sc->nSlots = sc->nEntries;
    rp->v[RP_DT_NEXT] = sc->deltaT;
    rp->v[RP_N_SORTED] = (double)sc->nSorted;
    rp->v[RP_N_CLONED] = ncCounts ? (double)ncCounts[0] : 0.0;
    rp->v[RP_N_ELIMINATED] = ncCounts ? (double)ncCounts[1] : 0.0;
    const int nCl = ncCounts ? ncCounts[0] : 0;
    sc->nTail = nCl;
    sc->nInjStart = nCl;
    sc->nextId += (uint64_t)nCl * (uint64_t)nRanksForId;
    sc->anyLost = rp->v[RP_K1SUM + ACC_N_LOST] > 0 ? 1 : 0;
    sc->nLostTotal += (long long)rp->v[RP_K1SUM + ACC_N_LOST];
}

// last kernel of a step: the finished report goes into its slot of the ring that the host reads once per batch of steps
__global__ void k_report_out(StepScalars *sc, DevReport *rp, DevReport *ring, int ringSize) {
    const int slot = sc->ringIdx % ringSize;
    if (threadIdx.x == 0) rp->v[RP_OVERFLOW] = (double)sc->overflow;
    __syncthreads();
    if (threadIdx.x < 64) ring[slot].v[threadIdx.x] = rp->v[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) sc->ringIdx = sc->ringIdx + 1;
}

__global__ void k_rng_probe(RngKey key, int n, const uint64_t *entity, uint32_t step, uint32_t purpose, uint32_t sub, double *out6) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *o = out6 + 6 * (size_t)i;
    rngUniform2(key, entity[i], step, purpose, sub, o[0], o[1]);
    rngNormal4(key, entity[i], step, purpose, sub, o[2], o[3], o[4], o[5]);
}

// ===========================================================================
// host side
// ===========================================================================
// exclusive scan of n ints (+ total): one CTA for small n, three kernels otherwise
static void scanInts(Cloud &c, int n, const int *in, int *out, int *total) {
    if (n <= 2 * SCAN_TILE) {
        LAUNCH(&c, k_scan_single, 1, 1024, 0, n, in, out, total);
        return;
    }
    const int nTiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if ((int)c.scanTileSums.n < nTiles) { c.scanTileSums.alloc(nTiles); c.scanTileOffs.alloc(nTiles); }
    LAUNCH(&c, k_scan_tile_sums, nTiles, 256, 0, n, in, c.scanTileSums.p);
    LAUNCH(&c, k_scan_single, 1, 1024, 0, nTiles, c.scanTileSums.p, c.scanTileOffs.p, total);
    LAUNCH(&c, k_scan_tiles, nTiles, 1024, 0, n, in, c.scanTileOffs.p, out);
}

static RngKey makeKey(uint64_t seed) {
    RngKey k;
    k.k0 = (uint32_t)(seed & 0xffffffffu);
    k.k1 = (uint32_t)(seed >> 32);
    return k;
}

static FlameletTables tablesOf(const Cloud &c) {
    FlameletTables f;
    f.chi = c.flChi.p; f.z = c.flZ.p; f.fields = c.flFields.p; f.nChi = c.nChi; f.nZ = c.nZ; f.cEq = c.flCEq.p; f.cEqMax = c.cEqMax;
    return f;
}

extern "C" int pdfb200_rng_probe(int device, uint64_t seed, int64_t n, const uint64_t *entity, uint32_t step, uint32_t purpose,
                                 uint32_t sub, double *out6) {
    if (n <= 0) return PDFB200_OK;
    if (cudaSetDevice(device) != cudaSuccess) return PDFB200_ERR_CUDA;
    uint64_t *dE = nullptr;
    double *dO = nullptr;
    cudaError_t e = cudaMalloc((void **)&dE, n * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&dO, n * 6 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(dE, entity, n * sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_rng_probe<<<gridFor(n, 256), 256>>>(makeKey(seed), (int)n, dE, step, purpose, sub, dO);
        e = cudaMemcpy(out6, dO, n * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(dE); cudaFree(dO);
    return e == cudaSuccess ? PDFB200_OK : PDFB200_ERR_CUDA;
}

void Cloud::allocParticles() {
    if (!stateAllocs.empty()) return;
    const size_t cap = (size_t)capacity;
    auto allocD = [&](double *&p) { CUDA_CHECK(cudaMalloc((void **)&p, cap * sizeof(double))); stateAllocs.push_back(p); };
    auto allocI = [&](int *&p) { CUDA_CHECK(cudaMalloc((void **)&p, cap * sizeof(int))); stateAllocs.push_back(p); };
    for (int b = 0; b < 2; ++b) {
        PState &s = S[b];
        allocD(s.px); allocD(s.py); allocD(s.pz); allocD(s.ux); allocD(s.uy); allocD(s.uz);
        allocD(s.m); allocD(s.omega); allocD(s.rho); allocD(s.eta);
        for (int k = 0; k < PDFB200_MAX_PHI; ++k) { s.phi[k] = nullptr; if (k < prm.nPhi) allocD(s.phi[k]); }
        allocI(s.cell); allocI(s.tet);
        double *idp; allocD(idp); s.id = reinterpret_cast<uint64_t *>(idp);   // 64-bit ids: a column as wide as a double
    }
    keysA.alloc(cap); keysB.alloc(cap); valsB.alloc(cap); iota.alloc(cap);
    // dynamic work distribution of the particle kernel (evolve_kernel.cuh); PDFB200_K1_DYNAMIC=0 selects the static
    // waves, PDFB200_K1_CHUNK the chunk size of a kernel specialised with -DK1_CHUNK=... (tuning experiments)
    if (const char *e = std::getenv("PDFB200_K1_DYNAMIC")) k1Dynamic = std::atoi(e);
    if (const char *e = std::getenv("PDFB200_K1_CHUNK")) k1Chunk = std::max(32, std::atoi(e));
    k1WorkCtr.alloc(1); k1WorkCtr.zero(stream);
    chunkSums.alloc(((size_t)cap / k1Chunk + 8) * 4 * K1_NC1); chunkSums.zero(stream);   // rows of scalars the case does not conserve stay zero
    chunkPart.alloc((size_t)numSMs * 4 * 4 * K1_NC1);
    LAUNCH(this, k_iota, gridFor((long long)cap, 256), 256, 0, iota.p, (int)cap);
    injPatchTail.alloc(cap);
    size_t tempBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keysA.p, keysB.p, iota.p, valsB.p, (int)cap, 0, keyBits + 1, stream);
    cubTemp.alloc(tempBytes + 256);
    cellStart.alloc(nCells); cellEnd.alloc(nCells); ncFlag.alloc(nCells); ncCount.alloc(nCells + 1); ncOffset.alloc(nCells + 1);
    scanTmp.alloc(8);
    csCount.alloc(nCells); csCursor.alloc(nCells); csStart.alloc(nCells); ncCounts.alloc(2);
    injCount.alloc(std::max(nBnd, 1)); injOffset.alloc(std::max(nBnd, 1) + 1); injSums.alloc((size_t)std::max(nBnd, 1) * K1_NC1);
    ncDelta.alloc((size_t)nCells * K1_NC1);
    if (axisym) cloneDelta.alloc((size_t)capacity * K1_NC1);
    Un.alloc(std::max(nBnd, 1)); Un.zero(stream);
    inletU.alloc((size_t)std::max(nBnd, 1) * 3); inletR.alloc((size_t)std::max(nBnd, 1) * 6); fwdTrans.alloc((size_t)std::max(nBnd, 1) * 9);
    probVec.alloc(std::max<size_t>(hFaceStart[nFaces] - hFaceStart[nInternalFaces], 1));
    const int maxGrid = numSMs * 16;
    k1PartSum.alloc((size_t)maxGrid * ACC_NSUM); k1PartMax.alloc((size_t)std::max(maxGrid * ACC_NMAX, numSMs * 64));
    cellPartSum.alloc((size_t)maxGrid * 8); finalSums.alloc(256);
    const size_t lostCap = (size_t)std::min<int64_t>(capacity, 65536);
    lostList.alloc(lostCap); lostSorted.alloc(lostCap); lostCount.alloc(1); lostCount.zero(stream);
    // nothing may be allocated while a step is being captured into a graph: scan scratch for the largest scan up front
    {
        const int nTiles = (std::max(nCells, nBnd) + SCAN_TILE - 1) / SCAN_TILE + 1;
        scanTileSums.alloc(nTiles); scanTileOffs.alloc(nTiles);
    }
    reportRing.alloc((size_t)REPORT_RING * 64);
    if (!hRing) CUDA_CHECK(cudaMallocHost((void **)&hRing, (size_t)REPORT_RING * 64 * sizeof(double)));
    if (evMarks.empty()) {
        evMarks.resize((size_t)REPORT_RING * 7);
        for (auto &e : evMarks) CUDA_CHECK(cudaEventCreate(&e));
    }
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Cloud::seedParticles() {
    dropGraphs();
    if (!haveMoments) throw std::logic_error("seed_particles: init_moments first");
    requireTables();
    allocParticles();
    const int Npc = prm.particlesPerCell;
    const int nLoc = (Npc - rank + nRanks - 1) / nRanks;
    const long long n = (long long)nCells * nLoc;
    if (n > capacity) throw std::length_error("seed_particles: capacity too small");
    cur = 0;
    prepareFields();
    LAUNCH(this, k_seed, gridFor(nCells, 64), 64, 0, dm, prm, S[cur], makeKey(prm.seed), rank, nRanks, rho_c.p, Ufv_c.p, kfv_c.p, Phi_c.p);
    LAUNCH(this, k_init_models, gridFor(n, 128), 128, 0, dm, prm, S[cur], (int)n, nodeMid.p, cellAux.p, cellAuxStride, scal.p, tablesOf(*this));
    StepScalars h = *hScal;
    h.nSorted = 0; h.nTail = (int)n; h.nInjStart = (int)n; h.nSlots = 0; h.anyLost = 0;
    h.nextId = (uint64_t)((long long)nCells * Npc + rank);
    *hScal = h;
    CUDA_CHECK(cudaMemcpyAsync(scal.p, hScal, sizeof(StepScalars), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    nParticlesHost = n; nSlotsHost = 0; sortedValid = false;
}

// labels handed in by the caller index the mesh tables without bounds checks in every kernel
__global__ void k_check_labels(DMesh M, int n, const int *cell, const int *tet, int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell[i];
    bool ok = c >= 0 && c < M.nCells;
    if (ok && tet) { const int t = tet[i]; ok = t >= M.cellTetStart[c] && t < M.cellTetStart[c + 1]; }
    if (!ok) atomicAdd(bad, 1);
}
// ids when the caller gives none: running index, interleaved over the ranks so that no two ranks share one
__global__ void k_default_ids(int n, uint64_t *id, int rank, int nRanks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) id[i] = (uint64_t)i * (uint64_t)nRanks + (uint64_t)rank;
}

void Cloud::uploadParticles(int64_t n, const pdfb200_particles &p) {
    dropGraphs();
    allocParticles();
    if (n > capacity) throw std::length_error("upload_particles: capacity too small");
    if (!p.pos || !p.U || !p.m || !p.Omega || !p.rho || !p.eta || !p.cell) throw std::invalid_argument("upload_particles: null column");
    cur = 0;
    PState &s = S[cur];
    std::vector<double> tmp((size_t)n);
    auto upStrided = [&](double *dst, const double *src, int comp) {
        for (int64_t i = 0; i < n; ++i) tmp[i] = src[3 * i + comp];
        CUDA_CHECK(cudaMemcpy(dst, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    };
    upStrided(s.px, p.pos, 0); upStrided(s.py, p.pos, 1); upStrided(s.pz, p.pos, 2);
    upStrided(s.ux, p.U, 0); upStrided(s.uy, p.U, 1); upStrided(s.uz, p.U, 2);
    auto up = [&](double *dst, const double *src) { CUDA_CHECK(cudaMemcpy(dst, src, n * sizeof(double), cudaMemcpyHostToDevice)); };
    up(s.m, p.m); up(s.omega, p.Omega); up(s.rho, p.rho); up(s.eta, p.eta);
    for (int k = 0; k < prm.nPhi; ++k) { if (!p.Phi[k]) throw std::invalid_argument("upload_particles: null scalar column"); up(s.phi[k], p.Phi[k]); }
    CUDA_CHECK(cudaMemcpy(s.cell, p.cell, n * sizeof(int), cudaMemcpyHostToDevice));
    // the next free id: above every id in use and congruent to the rank (ids advance in steps of nRanks)
    uint64_t idEnd = 0;
    if (p.id) {
        CUDA_CHECK(cudaMemcpy(s.id, p.id, n * sizeof(uint64_t), cudaMemcpyHostToDevice));
        for (int64_t i = 0; i < n; ++i) idEnd = std::max<uint64_t>(idEnd, (uint64_t)p.id[i] + 1);
    } else {
        if (n > 0) LAUNCH(this, k_default_ids, gridFor(n, 256), 256, 0, (int)n, s.id, rank, nRanks);
        idEnd = (uint64_t)n * (uint64_t)nRanks;
    }
    const uint64_t nextId64 = (idEnd + (uint64_t)nRanks - 1) / (uint64_t)nRanks * (uint64_t)nRanks + (uint64_t)rank;
    if (p.tet) CUDA_CHECK(cudaMemcpy(s.tet, p.tet, n * sizeof(int), cudaMemcpyHostToDevice));
    if (n > 0) {
        // labels are validated before any kernel indexes a mesh table with them (cell first: the locate kernel needs it)
        DBuf<int> bad; bad.alloc(1); bad.zero(stream);
        LAUNCH(this, k_check_labels, gridFor(n, 256), 256, 0, dm, (int)n, s.cell, p.tet ? s.tet : nullptr, bad.p);
        int hBad = 0;
        CUDA_CHECK(cudaMemcpyAsync(&hBad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        if (hBad) throw std::invalid_argument("upload_particles: " + std::to_string(hBad) + " particles carry a cell label outside the mesh or a tet label outside their cell");
        if (!p.tet) LAUNCH(this, k_locate, gridFor(n, 128), 128, 0, dm, (int)n, s);
    }
    StepScalars h = *hScal;
    h.nSorted = 0; h.nTail = (int)n; h.nInjStart = (int)n; h.nSlots = 0; h.anyLost = 0;
    h.nextId = nextId64;
    *hScal = h;
    CUDA_CHECK(cudaMemcpyAsync(scal.p, hScal, sizeof(StepScalars), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    nParticlesHost = n; nSlotsHost = 0; sortedValid = false;
}

// adds the share of the mass lost in the last step to the particles now (the particle kernel would do it at the
// start of the next step), so that a download shows the state the reference has after mcParticleCloud.C:1363-1369
__global__ void k_materialise_lost(int n, int nSorted, int nSlots, const uint32_t *perm, PState S, const double *lostShare, StepScalars *sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) sc->anyLost = 0;
    if (i >= n) return;
    const int slot = i < nSorted ? (int)perm[i] : nSlots + (i - nSorted);
    const int c = S.cell[slot];
    if (c >= 0) S.m[slot] += lostShare[c];
}

// gathers column `src` through idx (or identity for the tail) into `dst`
__global__ void k_gather_d(int n, int nSorted, int nSlots, const uint32_t *perm, const double *src, double *dst, const int *cell, int *alive) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = i < nSorted ? (int)perm[i] : nSlots + (i - nSorted);
    dst[i] = src[slot];
    if (alive) alive[i] = cell[slot] >= 0;
}
__global__ void k_gather_i(int n, int nSorted, int nSlots, const uint32_t *perm, const int *src, int *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = i < nSorted ? (int)perm[i] : nSlots + (i - nSorted);
    dst[i] = src[slot];
}

__global__ void k_gather_u64(int n, int nSorted, int nSlots, const uint32_t *perm, const uint64_t *src, uint64_t *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = i < nSorted ? (int)perm[i] : nSlots + (i - nSorted);
    dst[i] = src[slot];
}

void Cloud::downloadParticles(int64_t maxn, pdfb200_particles &out, int64_t *nOut) {
    CUDA_CHECK(cudaMemcpy(hScal, scal.p, sizeof(StepScalars), cudaMemcpyDeviceToHost));
    const int nSorted = hScal->nSorted, nTail = hScal->nTail, nSlots = hScal->nSlots;
    const int n = nSorted + nTail;
    if (hScal->anyLost && n > 0) {
        LAUNCH(this, k_materialise_lost, gridFor(n, 256), 256, 0, n, nSorted, nSlots, valsB.p, S[cur], lostShare.p, scal.p);
        hScal->anyLost = 0;
    }
    // logical order: sorted particles, then the tail; eliminated ones (cell < 0) are skipped
    DBuf<double> d; d.alloc(std::max(n, 1));
    DBuf<int> di, alive; di.alloc(std::max(n, 1)); alive.alloc(std::max(n, 1));
    std::vector<double> h(n);
    std::vector<int> hi(n), hal(n);
    const uint32_t *perm = valsB.p;
    const PState &s = S[cur];
    const int B = 256, G = gridFor(n, B);
    auto col = [&](const double *src) {
        if (n) { LAUNCH(this, k_gather_d, G, B, 0, n, nSorted, nSlots, perm, src, d.p, s.cell, alive.p); d.download(h.data(), n, stream); alive.download(hal.data(), n, stream); }
        CUDA_CHECK(cudaStreamSynchronize(stream));
    };
    auto coli = [&](const int *src) {
        if (n) { LAUNCH(this, k_gather_i, G, B, 0, n, nSorted, nSlots, perm, src, di.p); di.download(hi.data(), n, stream); }
        CUDA_CHECK(cudaStreamSynchronize(stream));
    };
    col(s.m);
    int64_t na = 0;
    for (int i = 0; i < n; ++i) na += hal[i] != 0;
    if (na > maxn) throw std::length_error("download_particles: buffer too small");
    std::vector<int> aliveIdx; aliveIdx.reserve(na);
    for (int i = 0; i < n; ++i) if (hal[i]) aliveIdx.push_back(i);
    auto put = [&](double *dst, int stride, int comp) { if (dst) for (int64_t k = 0; k < na; ++k) dst[stride * k + comp] = h[aliveIdx[k]]; };
    put(out.m, 1, 0);
    col(s.px); put(out.pos, 3, 0); col(s.py); put(out.pos, 3, 1); col(s.pz); put(out.pos, 3, 2);
    col(s.ux); put(out.U, 3, 0); col(s.uy); put(out.U, 3, 1); col(s.uz); put(out.U, 3, 2);
    col(s.omega); put(out.Omega, 1, 0); col(s.rho); put(out.rho, 1, 0); col(s.eta); put(out.eta, 1, 0);
    for (int k = 0; k < prm.nPhi; ++k) { col(s.phi[k]); put(out.Phi[k], 1, 0); }
    coli(s.cell); if (out.cell) for (int64_t k = 0; k < na; ++k) out.cell[k] = hi[aliveIdx[k]];
    coli(s.tet); if (out.tet) for (int64_t k = 0; k < na; ++k) out.tet[k] = hi[aliveIdx[k]];
    if (out.id) {
        // 64-bit ids: gathered as 8-byte words into the double staging buffer, copied bit for bit
        if (n) { LAUNCH(this, k_gather_u64, G, B, 0, n, nSorted, nSlots, perm, s.id, reinterpret_cast<uint64_t *>(d.p)); d.download(h.data(), n, stream); }
        CUDA_CHECK(cudaStreamSynchronize(stream));
        for (int64_t k = 0; k < na; ++k) std::memcpy(&out.id[k], &h[aliveIdx[k]], sizeof(uint64_t));
    }
    *nOut = na;
}

// ---------------------------------------------------------------------------
// one time step
// ---------------------------------------------------------------------------
template <int N>
static void launchMoments(Cloud &c, const PState &S, const uint32_t *perm) {
    const int B = 256, warps = B / 32;
    const int G = std::min(gridFor(c.nCells, warps), c.numSMs * 16);
    LAUNCH(&c, k_moments<N>, G, B, 0, c.dm, S, perm, c.cellStart.p, c.cellEnd.p, c.momentBlock.p, c.nMom);
}
// segment ordering of the counting sort + moments in one kernel (k_order_moments)
template <int N>
static void launchOrderMoments(Cloud &c, const PState &S) {
    const int B = 256, warps = B / 32;
    const int G = std::min(gridFor(c.nCells, warps), c.numSMs * 16);
    LAUNCH(&c, k_order_moments<N>, G, B, 0, c.dm, S, c.csStart.p, c.csCount.p, c.cellStart.p, c.cellEnd.p, c.valsB.p, c.keysB.p, c.momentBlock.p, c.nMom);
}
template <template <int> class F, class... Args>
static void dispatchNPhi(int nPhi, Args &&...args) {
    switch (nPhi) {
    case 0: F<0>::run(args...); break;
    case 1: F<1>::run(args...); break;
    case 2: F<2>::run(args...); break;
    case 3: F<3>::run(args...); break;
    case 4: F<4>::run(args...); break;
    case 5: F<5>::run(args...); break;
    default: F<6>::run(args...); break;
    }
}
template <int N> struct MomentsLauncher { static void run(Cloud &c, const PState &S, const uint32_t *perm) { launchMoments<N>(c, S, perm); } };
template <int N> struct OrderMomentsLauncher { static void run(Cloud &c, const PState &S) { launchOrderMoments<N>(c, S); } };

// Everything one step launches, in order, on the library's stream. Nothing in here waits for the device or reads a
// device value on the host (entry counts, time step, overflow flags all stay on the device; the step's report goes
// into a ring, see evolve()), so a step can be issued behind the previous one - and captured into a CUDA graph.
// `ev`: seven events that bracket the six regions of pdfb200_profile, or nullptr (graph capture).
void Cloud::enqueueStep(void *k1Jit, size_t k1JitSmem, int k1Blocks, cudaEvent_t *ev) {
    const RngKey key = makeKey(prm.seed);
    DevReport *dRep = reinterpret_cast<DevReport *>(finalSums.p + 128);
    DevReport *dRing = reinterpret_cast<DevReport *>(reportRing.p);
    auto mark = [&](int r) { if (ev) CUDA_CHECK(cudaEventRecord(ev[r], stream)); };
    void (*k1)(const K1Args) = prm.nPhi <= 1 ? k_evolve<1> : prm.nPhi == 2 ? k_evolve<2> : prm.nPhi == 3 ? k_evolve<3> : k_evolve<PDFB200_MAX_PHI>;
    const long long cap = capacity;

    // ---- cell-side preparation (all updateInternals) ----
    mark(0);
    prepareFields();
    if (hasOpen) LAUNCH(this, k_open_un, gridFor(nBnd, 128), 128, 0, dm, Ufv_b.p, Un.p);
    if (hasInlet) {
        if (!inletCached) {
            LAUNCH(this, k_inlet_cache, gridFor(nBnd, 128), 128, 0, dm, prm.k0, Ufv_b.p, Tau_b.p, fwdTrans.p, probVec.p, inletU.p, inletR.p);
            inletCached = true;
        }
        for (int pass = 0; pass < 2; ++pass) {
            LAUNCH(this, k_inject, std::min(gridFor((long long)nBnd * 32, 128), numSMs * 16), 128, 0, dm, prm, pass, S[cur], key, scal.p, Un.p, rho_b.p, Phi_b.p, inletU.p,
                   inletR.p, fwdTrans.p, probVec.p, nodeMid.p, injCount.p, injOffset.p, injPatchTail.p, injSums.p, nRanks, rank, cap);
            if (pass == 0) scanInts(*this, nBnd, injCount.p, injOffset.p, scanTmp.p);
        }
        LAUNCH(this, k_inject_finish, 1, 32, 0, nBnd, scanTmp.p, scal.p, cap, nRanks);
        const int gI = std::min(gridFor(nBnd, 256), numSMs * 2);
        LAUNCH(this, k_colsum_nc1, gI, 256, 0, nBnd, injSums.p, k1PartSum.p);
        LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gI, K1_NC1, k1PartSum.p, dRep->v + RP_INJ, 0);
    } else {
        CUDA_CHECK(cudaMemsetAsync(dRep->v + RP_INJ, 0, K1_NC1 * sizeof(double), stream));
    }
    // entry count of the step (device side), empty lost-particle list
    LAUNCH(this, k_step_begin, 1, 32, 0, scal.p, dRep, lostCount.p, k1WorkCtr.p);

    // ---- the fused particle kernel: persistent grid, the kernel reads the entry count itself ----
    K1Args A;
    A.M = dm; A.P = prm; A.in = S[cur]; A.out = S[cur ^ 1];
    A.perm = valsB.p; A.sc = scal.p;
    A.nodeMid = nodeMid.p; A.nodePC = nodePC.p; A.cellAux = cellAux.p; A.auxStride = cellAuxStride; A.pcScale = pcScale;
    A.Un = Un.p; A.lostShare = lostShare.p; A.lostList = lostList.p; A.lostCount = lostCount.p; A.lostCap = (unsigned)lostList.n;
    A.injPatchTail = injPatchTail.p;
    A.flChi = flChi.p; A.flZ = flZ.p; A.flFields = flFields.p; A.nChi = nChi; A.nZ = nZ; A.flCEq = flCEq.p; A.cEqMax = cEqMax;
    A.key = key; A.partSum = k1PartSum.p; A.partMax = k1PartMax.p; A.hasOpen = hasOpen;
    A.ctaClock = ctaClock.n ? ctaClock.p : nullptr;
    A.workCtr = k1Dynamic ? k1WorkCtr.p : nullptr; A.chunkSums = chunkSums.p;
    const int gridK1 = (int)std::max<long long>(1, std::min<long long>(gridFor(cap, K1_THREADS), (long long)numSMs * k1Blocks));
    // the lost mass of the previous step has been handed out (lostShare): its cells start empty again
    CUDA_CHECK(cudaMemsetAsync(lostMass.p, 0, nCells * sizeof(double), stream));
    if (ctaClock.n) ctaClockGrid = std::min<long long>(gridK1, (long long)ctaClock.n / 3);
    mark(1);
    if (k1Jit) {
        void *kargs[] = {&A};
        CUDA_CHECK(cudaLaunchKernel(k1Jit, dim3(gridK1), dim3(K1_THREADS), kargs, k1JitSmem, stream));
        ++launches;
    } else {
        LAUNCH(this, k1, gridK1, K1_THREADS, K1_SMEM_BYTES, A);
    }
    mark(2);
    LAUNCH(this, k_lost_reduce, 1, 1024, 0, lostList.p, lostSorted.p, lostCount.p, (unsigned)lostList.n, lostMass.p, scal.p);
    LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gridK1, ACC_NSUM, k1PartSum.p, dRep->v + RP_K1SUM, 0);
    LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gridK1, ACC_NMAX, k1PartMax.p, dRep->v + RP_K1MAX, 1);
    if (k1Dynamic) {
        // the mass sums come chunk by chunk (fixed order inside a chunk, fixed order over the chunks)
        const int gC = (int)std::max<long long>(1, std::min<long long>(numSMs * 4, (cap / k1Chunk + 63) / 64));   // >= 64 chunks per CTA
        LAUNCH(this, k_chunk_fold, gC, 256, 0, scal.p, k1Chunk, chunkSums.p, chunkPart.p);
        LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gC, 4 * K1_NC1, chunkPart.p, dRep->v + RP_K1SUM, 0);
    }
    cur ^= 1;

    // ---- sort by cell + segment bounds ----
    static const bool fuseEnv = [] { const char *e = std::getenv("PDFB200_FUSE_ORDER"); return !e || std::atoi(e) != 0; }();
    const bool fusedOrder = !radixPath && fuseEnv && !(prm.mixingModel == PDFB200_MIXING_MODCURL && prm.nMixed > 0);
    if (!radixPath) {
        // counting sort by cell (see k_cs_*): same permutation as the stable radix sort
        CUDA_CHECK(cudaMemsetAsync(csCount.p, 0, nCells * sizeof(int), stream));
        CUDA_CHECK(cudaMemsetAsync(csCursor.p, 0, nCells * sizeof(int), stream));
        const int gS = std::min(gridFor(cap, 256), numSMs * 16);
        LAUNCH(this, k_cs_hist, gS, 256, 0, scal.p, S[cur].cell, cellRank.p, keysA.p, csCount.p);
        scanInts(*this, nCells, csCount.p, csStart.p, scanTmp.p + 2);
        LAUNCH(this, k_cs_scatter, gS, 256, 0, scal.p, keysA.p, csStart.p, csCursor.p, valsB.p);
        mark(3);
        // the segments are put into their canonical order by the moments kernel itself (k_order_moments) unless the
        // pair mixing has to run on the ordered segments first
        if (!fusedOrder)
            LAUNCH(this, k_cs_order, std::min(gridFor(nCells, 8), numSMs * 16), 256, 0, nCells, csStart.p, csCount.p, cellOfRank.p, cellStart.p,
                   cellEnd.p, valsB.p, keysB.p);
        LAUNCH(this, k_cs_finish, 1, 32, 0, scanTmp.p + 2, scal.p);
    } else {
        // radix sort by (cell, tet): the library call needs the entry count on the host, so this path waits for it
        // (cells of 256 and more particles per rank only; evolve() issues one step per batch here)
        CUDA_CHECK(cudaMemcpyAsync(hScal, scal.p, sizeof(StepScalars), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        const int total = hScal->nEntries;
        size_t tempBytes = cubTemp.n;
        if (total > 0) {
            LAUNCH(this, k_make_keys, gridFor(total, 256), 256, 0, dm, total, S[cur].cell, S[cur].tet, keysA.p);
            CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cubTemp.p, tempBytes, keysA.p, keysB.p, iota.p, valsB.p, total, 0, keyBits + 1, stream));
        }
        mark(3);
        CUDA_CHECK(cudaMemsetAsync(cellStart.p, 0, nCells * sizeof(int), stream));
        CUDA_CHECK(cudaMemsetAsync(cellEnd.p, 0, nCells * sizeof(int), stream));
        if (total > 0) LAUNCH(this, k_cell_bounds, gridFor(total, 256), 256, 0, total, keysB.p, tetBits, cellOfRank.p, cellStart.p, cellEnd.p, scal.p);
        else CUDA_CHECK(cudaMemsetAsync(&scal.p->nSorted, 0, sizeof(int), stream));
    }
    // "modifiedCurl" pair mixing on the sorted order, before the moments are taken
    if (prm.mixingModel == PDFB200_MIXING_MODCURL && prm.nMixed > 0) {
        const int gM = std::min(gridFor(nCells, MIX_WARPS), numSMs * 8);
        LAUNCH(this, k_mix_curl, gM, MIX_WARPS * 32, 0, dm, prm, S[cur], valsB.p,
               cellStart.p, cellEnd.p, scal.p, key, keysA.p, valsA.p, cellPartSum.p);
        LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gM, K1_MAXCONS, cellPartSum.p, dRep->v + RP_MIXDELTA, 0);
    } else {
        CUDA_CHECK(cudaMemsetAsync(dRep->v + RP_MIXDELTA, 0, K1_MAXCONS * sizeof(double), stream));
    }
    if (fusedOrder) dispatchNPhi<OrderMomentsLauncher>(prm.nPhi, *this, S[cur]);
    else dispatchNPhi<MomentsLauncher>(prm.nPhi, *this, S[cur], (const uint32_t *)valsB.p);
    mark(4);
    // ---- cross-GPU sum of the moment block ----
    float arMs = 0;
    if (nRanks > 1) commAllreduceMoments(*this, &arMs);
    mark(5);

    // ---- time-averaging, mean fields, boundary conditions ----
    const int gF = std::min(gridFor(nCells, 256), numSMs * 8);
    // a pending asynchronous read-back of the cell fields must have left before they are overwritten (not while a
    // graph is being captured: evolve() orders the whole replay behind the copies instead)
    if (ev) orderAfterDownloads();
    fieldsFinalise(*this, cellPartSum.p, cellPartSum.p + (size_t)gF * 3, gF);
    LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gF, 3, cellPartSum.p, dRep->v + RP_CELLSUM, 0);
    LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gF, 2, cellPartSum.p + (size_t)gF * 3, dRep->v + RP_CELLMAX, 1);

    // ---- particle-number control ----
    CUDA_CHECK(cudaMemsetAsync(ncCounts.p, 0, 2 * sizeof(int), stream));
    const bool nc = prm.particleNumberControl != 0;
    if (nc) {
        CUDA_CHECK(cudaMemsetAsync(scanTmp.p + 3, 0, sizeof(int), stream));
        LAUNCH(this, k_nc_classify, gridFor(nCells, 256), 256, 0, dm, prm, PaNIC.p, cellStart.p, cellEnd.p, ncFlag.p, ncCount.p, nRanks,
               csCursor.p, scanTmp.p + 3);
        scanInts(*this, nCells, ncCount.p, ncOffset.p, scanTmp.p + 1);
        CUDA_CHECK(cudaMemsetAsync(ncDelta.p, 0, (size_t)nCells * K1_NC1 * sizeof(double), stream));
        // nSlots for the clones = this step's entry count
        LAUNCH(this, k_set_slots, 1, 32, 0, scal.p);
        const int gN = std::min(gridFor(nCells, 8), numSMs * 16);
        LAUNCH(this, k_nc_apply, gN, 256, 0, dm, prm, S[cur], valsB.p, cellStart.p, cellEnd.p, ncFlag.p, ncCount.p, ncOffset.p,
               PaNIC.p, scal.p, key, rank, nRanks, cap, ncDelta.p, ncCounts.p, keysA.p, csCursor.p, scanTmp.p + 3);
        // the clones: one thread per task (scanTmp[1] = number of tasks); keysA is free until the next step
        LAUNCH(this, k_nc_clone, numSMs * 8, 128, 0, dm, prm, S[cur], keysA.p, scanTmp.p + 1, scal.p, key, nRanks, cap,
               cloneDelta.p);
        // fold ncDelta over cells deterministically: per-CTA partials, then rows
        const int gR = std::min(gridFor(nCells, 256), numSMs * 4);
        LAUNCH(this, k_colsum_nc1, gR, 256, 0, nCells, ncDelta.p, k1PartSum.p);
        LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gR, K1_NC1, k1PartSum.p, dRep->v + RP_NCDELTA, 0);
        if (axisym) {
            LAUNCH(this, k_colsum_nc1, gR, 256, 0, 0, cloneDelta.p, k1PartSum.p, scanTmp.p + 1);
            LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gR, K1_NC1, k1PartSum.p, dRep->v + RP_CLONEDELTA, 0);
        } else {
            CUDA_CHECK(cudaMemsetAsync(dRep->v + RP_CLONEDELTA, 0, K1_NC1 * sizeof(double), stream));
        }
    } else {
        CUDA_CHECK(cudaMemsetAsync(dRep->v + RP_NCDELTA, 0, K1_NC1 * sizeof(double), stream));
        CUDA_CHECK(cudaMemsetAsync(dRep->v + RP_CLONEDELTA, 0, K1_NC1 * sizeof(double), stream));
    }
    // ---- lost mass shares ----
    {
        const int gL = std::min(gridFor(nCells, 256), numSMs * 4);
        LAUNCH(this, k_lost_share, gL, 256, 0, nCells, lostMass.p, PaNIC.p, cellStart.p, cellEnd.p, momentBlock.p, nMom, lostShare.p);
    }
    LAUNCH(this, k_step_end, 1, 32, 0, scal.p, dRep, prm, 0, nc ? ncCounts.p : nullptr, nRanks);
    {
        // after k_step_end: nSorted / nTail / nSlots describe the particles the next step starts from
        const int gB = numSMs * 4;
        LAUNCH(this, k_lost_balance, gB, 256, 0, dm, prm, S[cur], valsB.p, scal.p, lostCount.p, lostShare.p, k1PartSum.p);
        LAUNCH(this, k_fold_rows, 1, dim3(32, FOLD_Y), 0, gB, K1_NC1, k1PartSum.p, dRep->v + RP_LOSTSUM, 0);
    }
    LAUNCH(this, k_report_out, 1, 64, 0, scal.p, dRep, dRing, REPORT_RING);
    mark(6);
}

void Cloud::dropGraphs() {
    for (auto &g : stepGraph) { if (g) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(g)); g = nullptr; }
}

void Cloud::evolve(int nSteps, pdfb200_step_report *reports) {
    if (!haveMoments) throw std::logic_error("evolve: fields/moments not initialised");
    if (stateAllocs.empty()) throw std::logic_error("evolve: no particles (seed_particles or upload_particles first)");
    if (prm.nConserved > K1_MAXCONS) throw std::invalid_argument("evolve: at most 3 conserved scalars are supported");
    requireTables();
    if (prm.positionCorrection == PDFB200_POSCORR_EXTERNAL && !havePcExt)
        throw std::logic_error("evolve: positionCorrection external selected but no fields uploaded (upload_poscorr_fields)");
    if (nSteps <= 0) return;
    CUDA_CHECK(cudaEventRecord(evStart, stream));
    int k1Blocks = 0;
    // the scalars live in registers: instantiate the kernel for nPhi <= 1, 2, 3 and the general bound
    void (*k1)(const K1Args) = prm.nPhi <= 1 ? k_evolve<1> : prm.nPhi == 2 ? k_evolve<2> : prm.nPhi == 3 ? k_evolve<3> : k_evolve<PDFB200_MAX_PHI>;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k1Blocks, k1, K1_THREADS, K1_SMEM_BYTES));
    k1Blocks = std::max(1, std::min(k1Blocks, 16));
    // the kernel specialised for this cloud's parameters and mesh flags, when there is one (jit.cu)
    void *k1Jit = specialisedKernel();
    size_t k1JitSmem = K1_SMEM_BYTES;
    if (k1Jit) {
        k1Blocks = K1_JIT_MIN_BLOCKS;
        // the specialised kernel carries the sum rows of this case's conserved scalars only (evolve_kernel.cuh, K1_COMPACT_ACC)
        k1JitSmem = (size_t)K1_SMEM_ROWS(prm.nConserved) * K1_THREADS * sizeof(double);
        if (const char *x = std::getenv("PDFB200_K1_BLOCKS")) k1Blocks = std::max(1, std::atoi(x));   // with -DK1_MIN_BLOCKS in PDFB200_JIT_FLAGS
        if (const char *x = std::getenv("PDFB200_K1_SMEM")) k1JitSmem = (size_t)std::atol(x);          // with flags that change K1_SMEM_BYTES
    }
    static const bool forceRadix = [] { const char *e = std::getenv("PDFB200_SORT"); return e && std::string(e) == "radix"; }();
    radixPath = tetBits != 0 || forceRadix;
    if (prm.mixingModel == PDFB200_MIXING_MODCURL && valsA.n < (size_t)capacity) valsA.alloc(capacity);

    // Steps are issued in batches without waiting for the device in between: every step leaves its report in a ring on
    // the device, the host fetches the ring once per batch. One step per batch where the host has to see the entry
    // count of every step (the (cell, tet) radix sort of a library). With several ranks the all-reduce of every step is
    // timed by the step's own pair of events, so their steps are batched as well.
    // Small clouds are launch-bound (a step is ~50 kernels of a few microseconds): there the step is captured once into
    // a CUDA graph per state-buffer parity and replayed (PDFB200_GRAPH=0/1 overrides; no per-region timing in that mode).
    static const int graphEnv = [] { const char *e = std::getenv("PDFB200_GRAPH"); return e ? std::atoi(e) : -1; }();
    const bool mayGraph = !radixPath && nRanks == 1 && (!hasInlet || inletCached);
    const int gm = graphMode >= 0 ? graphMode : graphEnv;
    const bool useGraph = mayGraph && (gm >= 0 ? gm != 0 : capacity <= 4000000);
    const int batchMax = radixPath ? 1 : REPORT_RING;
    for (int done = 0; done < nSteps;) {
        const int nb = std::min(batchMax, nSteps - done);
        int zero = 0;
        CUDA_CHECK(cudaMemcpyAsync(&scal.p->ringIdx, &zero, sizeof(int), cudaMemcpyHostToDevice, stream));
        if (useGraph) orderAfterDownloads();
        for (int b = 0; b < nb; ++b) {
            if (useGraph) {
                void *&ge = stepGraph[cur];
                if (!ge) {
                    const long long launchesBefore = launches;
                    const int curBefore = cur;
                    cudaGraph_t g = nullptr;
                    CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                    try {
                        enqueueStep(k1Jit, k1JitSmem, k1Blocks, nullptr);
                    } catch (...) {
                        cudaStreamEndCapture(stream, &g);
                        if (g) cudaGraphDestroy(g);
                        cur = curBefore;
                        throw;
                    }
                    CUDA_CHECK(cudaStreamEndCapture(stream, &g));
                    cudaGraphExec_t exec = nullptr;
                    CUDA_CHECK(cudaGraphInstantiate(&exec, g, 0));
                    cudaGraphDestroy(g);
                    stepGraph[curBefore] = exec;
                    stepGraphLaunches[curBefore] = launches - launchesBefore;
                    launches = launchesBefore;
                    cur = curBefore;          // the capture has flipped it; the launch below does the step
                }
                CUDA_CHECK(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(stepGraph[cur]), stream));
                launches += stepGraphLaunches[cur];
                cur ^= 1;
            } else {
                enqueueStep(k1Jit, k1JitSmem, k1Blocks, evMarks.data() + (size_t)b * 7);
            }
        }
        // ---- the batch's reports ----
        CUDA_CHECK(cudaMemcpyAsync(hRing, reportRing.p, (size_t)nb * sizeof(DevReport), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaMemcpyAsync(hScal, scal.p, sizeof(StepScalars), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        for (int b = 0; b < nb; ++b) {
            const double *v = hRing + (size_t)b * 64;
            // the all-reduce of step b sits between marks 4 and 5 (several ranks never replay a graph)
            float arMs = 0;
            if (nRanks > 1 && !useGraph) CUDA_CHECK(cudaEventElapsedTime(&arMs, evMarks[(size_t)b * 7 + 4], evMarks[(size_t)b * 7 + 5]));
            if (!useGraph) {
                // per-region device times: prep+inject, evolve kernel, sort, bounds+moments, allreduce, finalise+control
                for (int r = 0; r < 6; ++r) {
                    float t = 0;
                    CUDA_CHECK(cudaEventElapsedTime(&t, evMarks[(size_t)b * 7 + r], evMarks[(size_t)b * 7 + r + 1]));
                    profMs[r] += t;
                }
                ++profSteps;
                profParticleSteps += v[RP_N_ENTRIES];
            }
            const int ov = (int)v[RP_OVERFLOW];
            if (ov == 2) throw std::runtime_error("evolve: more than " + std::to_string(lostList.n) + " particles lost in one step");

            if (ov) throw std::length_error("evolve: particle capacity exceeded");
            nParticlesHost = (int64_t)v[RP_N_SORTED] + (int64_t)v[RP_N_CLONED] - (int64_t)v[RP_N_ELIMINATED];
            nSlotsHost = (int64_t)v[RP_N_ENTRIES];
            if (reports) {
                pdfb200_step_report &r = reports[done + b];
                std::memset(&r, 0, sizeof(r));
                r.deltaT = v[RP_DT_USED]; r.deltaTNext = v[RP_DT_NEXT];
                r.rhoRes = v[RP_CELLSUM] / nCells; r.rhoErrMin = -v[RP_CELLMAX]; r.rhoErrMax = v[RP_CELLMAX + 1];
                r.totalMass = v[RP_CELLSUM + 1]; r.totalParticleMass = v[RP_CELLSUM + 2];
                r.UMax = std::max(0.0, v[RP_K1MAX + ACC_UMAX]); r.UcorrMax = std::max(0.0, v[RP_K1MAX + ACC_UCORRMAX]);
                r.etaMax = std::max(0.0, v[RP_K1MAX + ACC_ETAMAX]); r.etaMin = -v[RP_K1MAX + ACC_NEG_ETAMIN];
                r.CoMax = std::max(0.0, v[RP_K1MAX + ACC_COMAX]);
                for (int k = 0; k <= prm.nConserved; ++k) {
                    r.deltaMassInst[k] = v[RP_K1SUM + ACC_DM_MINUS + k] + v[RP_K1SUM + ACC_DM_PLUS + k] + (v[RP_NCDELTA + k] + v[RP_CLONEDELTA + k]) + v[RP_LOSTSUM + k];
                    r.massInInst[k] = v[RP_K1SUM + ACC_MASS_IN + k] + v[RP_INJ + k];
                    r.massOutInst[k] = v[RP_K1SUM + ACC_MASS_OUT + k];
                }
                for (int k = 1; k <= prm.nConserved; ++k) r.deltaMassInst[k] += v[RP_MIXDELTA + k - 1];
                r.nParticles = nParticlesHost;
                r.nInjected = (int64_t)v[RP_N_INJECTED]; r.nDeleted = (int64_t)v[RP_K1SUM + ACC_N_DELETED];
                r.nCloned = (int64_t)v[RP_N_CLONED]; r.nEliminated = (int64_t)v[RP_N_ELIMINATED]; r.nLost = (int64_t)v[RP_K1SUM + ACC_N_LOST];
                r.momentsAllreduceMs = arMs;
            }
        }
        deltaTHost = hScal->deltaT;
        stepHost = hScal->step;
        sortedValid = true;
        done += nb;
    }
    CUDA_CHECK(cudaEventRecord(evStop, stream));
    CUDA_CHECK(cudaEventSynchronize(evStop));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, evStart, evStop));
    lastEvolveMs = ms;
}
