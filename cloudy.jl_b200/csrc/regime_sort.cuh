// Stable counting sort of an ensemble's DATA into regime order (box-model states, structure of arrays).
//
// The thread-per-parcel kernel is fastest when the 32 parcels of a warp need the same series length and the same
// series / continued-fraction regime.  Round 1 walked the ensemble through a permutation (`KArgs::perm`): every 8-byte
// moment of a permuted parcel then costs a 64-byte DRAM access per slot array (measured on B200, 64 Mi parcels x 5 slots:
// 22.4 + 7.9 GB per launch for 5.4 GB of state; tiling the order down to 256 Ki parcels still moved 15 + 6.9 GB — L2 does
// not retain the tile, profiles/r02_sort_sweep.md).  Here the parcels themselves are moved into bucket order once and
// STAY in that order (parcels are independent, so their position in the device arrays is free): every later kernel reads
// and writes whole lines, and the order is refreshed every few time steps only (it is a scheduling hint; results do not
// depend on it).  `order[i]` = original parcel index of position i; uploads reset it, downloads undo it.
//
// Three kernels, deterministic (no order-dependent atomics), stable (bucket-internal order = previous position order):
//   regime_key_kernel (cloudy_b200.cu)   keys[p] (1 byte) + per-block histograms blockhist[bin][block] + totals[bin]
//   sort_scan_kernel                      exclusive prefix over the blocks of every occupied bin (one block per bin)
//   sort_scatter_kernel                   rank inside the block (per-warp match + prefix over warps), move all slots
#pragma once
#include "common.cuh"

namespace cloudy {

constexpr int SORT_THREADS = 256;
constexpr int SORT_KPT = 4;      // keys per thread and round (matches regime_key_kernel's KEY_PER_THREAD)
constexpr int SORT_ROUNDS = 4;   // rounds per block: a block owns 4096 consecutive positions
constexpr int SORT_BLOCK_PARCELS = SORT_THREADS * SORT_KPT * SORT_ROUNDS;

// one block per bin: blockhist[bin][0..nblocks) -> exclusive prefix in place
__global__ void __launch_bounds__(SORT_THREADS) sort_scan_kernel(unsigned int* __restrict__ blockhist, const unsigned int* __restrict__ totals, int nblocks) {
    const int bin = blockIdx.x;
    if (totals[bin] == 0) return;
    unsigned int* h = blockhist + (size_t)bin * nblocks;
    __shared__ unsigned int part[SORT_THREADS];
    const int per = (nblocks + SORT_THREADS - 1) / SORT_THREADS;
    const int lo = min(threadIdx.x * per, nblocks), hi = min(lo + per, nblocks);
    unsigned int s = 0;
    for (int i = lo; i < hi; ++i) s += h[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < SORT_THREADS; off <<= 1) {
        const unsigned int v = (threadIdx.x >= off) ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        const unsigned int c = h[i];
        h[i] = run;
        run += c;
    }
}

// moves every slot of every parcel to its sorted position and composes the order
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const unsigned char* __restrict__ keys, const unsigned int* __restrict__ blockoff,
                                                                  const unsigned int* __restrict__ totals, int nblocks, const double* __restrict__ in,
                                                                  double* __restrict__ out, long long stride, int nslots, long long n,
                                                                  const int* __restrict__ order_in, int* __restrict__ order_out) {
    __shared__ unsigned int binbase[256];  // exclusive prefix of the bin totals, then + this block's offset inside the bin
    __shared__ unsigned int wcnt[SORT_THREADS / 32][256];
    const unsigned int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const unsigned int own = totals[tid];
        binbase[tid] = own;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            const unsigned int v = (tid >= (unsigned)off) ? binbase[tid - off] : 0u;
            __syncthreads();
            binbase[tid] += v;
            __syncthreads();
        }
        const unsigned int excl = binbase[tid] - own;
        __syncthreads();
        binbase[tid] = excl + (own ? blockoff[(size_t)tid * nblocks + blockIdx.x] : 0u);
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) wcnt[w][tid] = 0;
    }
    __syncthreads();
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const long long p0 = ((long long)blockIdx.x * SORT_ROUNDS + r) * (SORT_KPT * SORT_THREADS);
        if (p0 >= n) break;  // block-uniform
        unsigned int key[SORT_KPT];
#pragma unroll
        for (int q = 0; q < SORT_KPT; ++q) {
            const long long p = p0 + q * SORT_THREADS + tid;
            key[q] = (p < n) ? keys[p] : 0u;
        }
#pragma unroll
        for (int q = 0; q < SORT_KPT; ++q) {
            const long long p = p0 + q * SORT_THREADS + tid;
            const bool live = p < n;
            // the parcel's data is requested before the ranking (three barriers) so that its latency is hidden
            double val[MAXSLOT];
#pragma unroll
            for (int s = 0; s < MAXSLOT; ++s) val[s] = (s < nslots && live) ? in[s * stride + p] : 0.0;
            const int oin = live ? (order_in ? order_in[p] : (int)p) : 0;
            const unsigned int act = __ballot_sync(0xffffffffu, live);
            unsigned int below = 0;
            if (live) {
                const unsigned int peers = __match_any_sync(act, key[q]);
                below = __popc(peers & ((1u << lane) - 1u));
                if (below == 0) wcnt[warp][key[q]] = __popc(peers);  // the group's lowest lane
            }
            __syncthreads();
            unsigned int dst = 0;
            if (live) {
                dst = binbase[key[q]] + below;
                for (unsigned int w = 0; w < warp; ++w) dst += wcnt[w][key[q]];
            }
            __syncthreads();
            {   // thread `tid` owns bin `tid`: advance the running base, clear the per-warp counts
                unsigned int add = 0;
#pragma unroll
                for (int w = 0; w < SORT_THREADS / 32; ++w) { add += wcnt[w][tid]; wcnt[w][tid] = 0; }
                binbase[tid] += add;
            }
            __syncthreads();
            if (live) {
                order_out[dst] = oin;
#pragma unroll
                for (int s = 0; s < MAXSLOT; ++s)
                    if (s < nslots) out[s * stride + dst] = val[s];
            }
        }
    }
}

// device AoS staging <-> SoA through an order (download / upload of regime-ordered states)
__global__ void soa_to_aos_ordered_kernel(const double* __restrict__ soa, double* __restrict__ aos, long long n, int nslots, long long stride,
                                          const int* __restrict__ order) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = n * nslots;
    for (; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / nslots;
        const int s = (int)(i % nslots);
        aos[(long long)order[p] * nslots + s] = soa[s * stride + p];
    }
}

}  // namespace cloudy
