// walk_probe.cu — measurement aid, not part of the step: a kernel that does nothing but the tet
// walk of mcParticle::move (the loop of evolve_kernel.cuh, without boundaries handling beyond
// "stop at a boundary face") for every particle of the cloud from its current position over
// eta*deltaT along its velocity. It answers one design question for the next round: how fast is
// the walk alone when it is not fused with the models, i.e. with few live registers and many
// resident warps (profiles/README.md). Exported as pdfb200_walk_probe (not in the public header).
#include "cloud.hpp"
#include "evolve_kernel.cuh"

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_walk_probe(DMesh M, PState S, const uint32_t *perm, int n, const StepScalars *sc,
                                                           int *outTet, unsigned long long *visits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int nv = 0;
    if (i < n) {
        const int slot = (int)perm[i];
        if (S.cell[slot] >= 0) {
            const V3 x0 = mk(S.px[slot], S.py[slot], S.pz[slot]);
            const V3 U = mk(S.ux[slot], S.uy[slot], S.uz[slot]);
            int tet = S.tet[slot];
            const double dt = S.eta[slot] * sc->deltaT;
            const V3 dx = (x0 + dt * U) - x0;
            TetL T;
            loadTet(M.tets, tet, T);
            int entry = -1;
            nv = 1;
            for (;;) {
                const V3 r = x0 - mk(T.cx, T.cy, T.cz);
                double p0, p1, p2, p3;
                baryAt(T, r, p0, p1, p2, p3);
                const double q0 = dotf(T.g0, T.g1, T.g2, dx), q1 = dotf(T.g3, T.g4, T.g5, dx), q2 = dotf(T.g6, T.g7, T.g8, dx);
                const double q3 = -((q0 + q1) + q2);
                const bool c0 = (entry != 0) & (q0 < 0) & (p0 + q0 < 0);
                const bool c1 = (entry != 1) & (q1 < 0) & (p1 + q1 < 0);
                const bool c2 = (entry != 2) & (q2 < 0) & (p2 + q2 < 0);
                const bool c3 = (entry != 3) & (q3 < 0) & (p3 + q3 < 0);
                int best = c0 ? 0 : -1;
                double pb = c0 ? p0 : 0.0, qb = c0 ? q0 : 0.0;
                const bool t1 = c1 & ((best < 0) | (p1 * qb > pb * q1));
                best = t1 ? 1 : best; pb = t1 ? p1 : pb; qb = t1 ? q1 : qb;
                const bool t2 = c2 & ((best < 0) | (p2 * qb > pb * q2));
                best = t2 ? 2 : best; pb = t2 ? p2 : pb; qb = t2 ? q2 : qb;
                const bool t3 = c3 & ((best < 0) | (p3 * qb > pb * q3));
                best = t3 ? 3 : best;
                if (best < 0 || nv > 1000) break;
                if (best == 3 && T.n3 < 0) break;   // boundary face: the probe stops here
                int nb = (best == 0) ? T.n0 : T.n1;
                nb = (best == 2) ? T.n2 : nb;
                tet = (best == 3) ? T.n3 : nb;
                loadTet(M.tets, tet, T);
                entry = best ^ (((best == 1) | (best == 2)) ? 3 : 0);
                ++nv;
            }
            outTet[i] = tet;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    if ((threadIdx.x & 31) == 0 && nv > 0) atomicAdd(visits, (unsigned long long)nv);
}

// minBlocks: CTAs of 128 threads per SM the kernel is compiled for (3..8). Returns the device time of
// `repeat` launches (ms) and the number of tet visits of one launch.
extern "C" int pdfb200_walk_probe(pdfb200_cloud *cloud, int minBlocks, int repeat, double *ms, long long *visitsOut) {
    Cloud &c = *reinterpret_cast<Cloud *>(cloud);
    cudaSetDevice(c.device);
    const int n = (int)c.nSlotsHost;
    if (n <= 0 || !c.sortedValid) return PDFB200_ERR_STATE;
    DBuf<unsigned long long> visits; visits.alloc(1);
    DBuf<int> outTet; outTet.alloc(n);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = (n + 127) / 128;
    auto launch = [&](void) {
        switch (minBlocks) {
        case 3: k_walk_probe<3><<<grid, 128, 0, c.stream>>>(c.dm, c.S[c.cur], c.valsB.p, n, c.scal.p, outTet.p, visits.p); break;
        case 4: k_walk_probe<4><<<grid, 128, 0, c.stream>>>(c.dm, c.S[c.cur], c.valsB.p, n, c.scal.p, outTet.p, visits.p); break;
        case 5: k_walk_probe<5><<<grid, 128, 0, c.stream>>>(c.dm, c.S[c.cur], c.valsB.p, n, c.scal.p, outTet.p, visits.p); break;
        case 6: k_walk_probe<6><<<grid, 128, 0, c.stream>>>(c.dm, c.S[c.cur], c.valsB.p, n, c.scal.p, outTet.p, visits.p); break;
        default: k_walk_probe<8><<<grid, 128, 0, c.stream>>>(c.dm, c.S[c.cur], c.valsB.p, n, c.scal.p, outTet.p, visits.p); break;
        }
    };
    visits.zero(c.stream);
    launch();   // warm-up + visit count
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, visits.p, sizeof h, cudaMemcpyDeviceToHost, c.stream);
    cudaStreamSynchronize(c.stream);
    cudaEventRecord(a, c.stream);
    for (int r = 0; r < repeat; ++r) launch();
    cudaEventRecord(b, c.stream);
    cudaEventSynchronize(b);
    float t = 0;
    cudaEventElapsedTime(&t, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms = t;
    *visitsOut = (long long)h;
    return cudaGetLastError() == cudaSuccess ? PDFB200_OK : PDFB200_ERR_CUDA;
}
