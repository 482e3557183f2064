// k1_layout.h — launch shape and accumulator layout of the particle kernel, shared by the host side
// (cloud.hpp) and by the kernel source, which is also compiled at run time (jit.cu) where no host
// header is available.
#pragma once

// layout of the particle kernel's per-thread accumulators
#ifndef K1_THREADS
#define K1_THREADS 128
#endif
// 1: the warps of the particle kernel claim chunks of entries from a counter; 0: static waves (evolve_kernel.cuh)
#ifndef K1_DYNAMIC
#define K1_DYNAMIC 1
#endif
// entries a warp claims at a time under the dynamic work distribution of the particle kernel (a multiple of 32)
#ifndef K1_CHUNK
#define K1_CHUNK 64
#endif
#define K1_JIT_MIN_BLOCKS 5   // CTAs per SM the run-time specialised kernel is compiled for (96 registers; measured best of 3..6)
#ifndef K1_MAXCONS
#define K1_MAXCONS 3
#endif
#define K1_NC1 (1 + K1_MAXCONS)
// thread-private shared-memory accumulators
#define ACC_DM_MINUS 0
#define ACC_DM_PLUS (ACC_DM_MINUS + K1_NC1)
#define ACC_MASS_IN (ACC_DM_PLUS + K1_NC1)
#define ACC_MASS_OUT (ACC_MASS_IN + K1_NC1)
#define ACC_N_DELETED (ACC_MASS_OUT + K1_NC1)
#define ACC_N_LOST (ACC_N_DELETED + 1)
#define ACC_N_ALIVE (ACC_N_LOST + 1)
#define ACC_NSUM (ACC_N_ALIVE + 1)
#define ACC_UMAX 0
#define ACC_UCORRMAX 1
#define ACC_ETAMAX 2
#define ACC_NEG_ETAMIN 3
#define ACC_COMAX 4
#define ACC_NMAX 5
// dynamic shared memory of the particle kernel: thread-private rows of K1_THREADS doubles - the sums (with the rows of
// nCons conserved scalars), the maxima, and K1_NPARK parking rows (the correction velocity and the Courant number,
// which must survive the tet walks without occupying registers)
#define K1_NPARK 4
#define K1_SMEM_ROWS(nCons) (4 * (1 + (nCons)) + 3 + ACC_NMAX + K1_NPARK)

