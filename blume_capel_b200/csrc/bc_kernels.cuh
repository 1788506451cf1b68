// bc_kernels.cuh -- hand-written sm_100a kernels of the checkerboard Metropolis half-sweep.
//
// Data layout in HBM: two compact per-colour arrays of spin codes c = s+1 in {0,1,2}, one
// byte per site, [replica][y][k] with k = x>>1 (site x = 2k + ((y+colour)&1)), row stride Wc.
// One 128-bit load = 16 same-colour sites of one row.  For a site of colour C in row y with
// p = (y+C)&1 its neighbours in the other colour's array O are O[y-1][k], O[y+1][k], O[y][k]
// and O[y][k+1] (p = 1) or O[y][k-1] (p = 0).
//
// Per site the kernels do what /root/reference/main.cpp:89-112 does per attempt, with the
// double-precision exp() replaced by the integer table of the spec.
//
// This header is the umbrella the engine includes; the kernels live in
//   bc_update.cuh          the per-row update (Philox draws, packed compares, pair table) shared by all fast kernels
//   bc_sweep_stream.cuh    streaming kernel, group launch (several lattice sizes), experimental persistent kernel
//   bc_sweep_wave.cuh      experimental wavefront kernel
//   bc_sweep_resident.cuh  shared-memory-resident kernel for small lattices
//   bc_sweep_generic.cuh   one thread per site, any L
//   bc_util_kernels.cuh    layout conversion, init, pair tables, slab halos, stand-alone measurement
//   bc_cluster_tile.cuh    bit-parallel labelling of one 64 x 64 tile by one warp (host + device phases)
//   bc_cluster_kernels.cuh Swendsen-Wang cluster move, cluster labels and cluster-size moments
//   bc_cluster_sort.cuh    cluster member lists (radix sort by label), gap-size histogram, dump-writer lists
#pragma once
#include "bc_common.cuh"
#include "bc_update.cuh"
#include "bc_sweep_stream.cuh"
#ifdef BC_EXPERIMENTAL
#include "bc_sweep_wave.cuh"  // measured negative result, kept as a record
#endif
#include "bc_sweep_resident.cuh"
#include "bc_sweep_generic.cuh"
#include "bc_util_kernels.cuh"
#include "bc_cluster_kernels.cuh"
#include "bc_cluster_sort.cuh"
