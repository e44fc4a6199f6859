// Probe: how many clusters of size CS (1 CTA per SM, 213 KB dynamic smem, 256 threads) can be co-resident on this GPU,
// and the latency of cluster.sync / a DSMEM round trip / the atomic-counter grid barrier with 128 CTAs in clusters.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void dummy(int* out) {
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0 && out) out[blockIdx.x] = sm[0];
}

__global__ void lat(long long* out, unsigned* ctr) {
    extern __shared__ unsigned char sm[];
    cg::cluster_group cl = cg::this_cluster();
    volatile float* mine = reinterpret_cast<volatile float*>(sm);
    mine[threadIdx.x] = (float)threadIdx.x;
    cl.sync();
    long long t0 = clock64();
    for (int i = 0; i < 16; ++i) cl.sync();
    long long t1 = clock64();
    // DSMEM dependent loads
    const float* peer = cl.map_shared_rank(reinterpret_cast<const float*>(sm), (cl.block_rank() + 1) % cl.num_blocks());
    float acc = 0.f;
    int idx = threadIdx.x;
    for (int i = 0; i < 16; ++i) {
        acc += peer[idx];
        idx = ((int)acc + i) & 255;
    }
    long long t2 = clock64();
    // grid barrier x16
    unsigned gen = 0;
    const unsigned nb = gridDim.x;
    for (int i = 0; i < 16; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            ++gen;
            while ((int)(*reinterpret_cast<volatile unsigned*>(ctr) - gen * nb) < 0) {
            }
            __threadfence();
        }
        __syncthreads();
    }
    long long t3 = clock64();
    // hierarchical: cluster.sync, one CTA per cluster arrives on the counter, cluster.sync
    unsigned* ctr2 = ctr + 32;
    const unsigned ncl = gridDim.x / cl.num_blocks();
    gen = 0;
    for (int i = 0; i < 16; ++i) {
        __threadfence();
        cl.sync();
        if (cl.block_rank() == 0 && threadIdx.x == 0) {
            atomicAdd(ctr2, 1u);
            ++gen;
            while ((int)(*reinterpret_cast<volatile unsigned*>(ctr2) - gen * ncl) < 0) {
            }
            __threadfence();
        }
        cl.sync();
    }
    long long t4 = clock64();
    if (threadIdx.x == 0) {
        out[blockIdx.x * 4 + 0] = (t1 - t0) / 16;
        out[blockIdx.x * 4 + 1] = (t2 - t1) / 16;
        out[blockIdx.x * 4 + 2] = (t3 - t2) / 16;
        out[blockIdx.x * 4 + 3] = (t4 - t3) / 16;
    }
    if (acc == 12345.f) out[0] = 0;
    cl.sync();
}

int main() {
    const int smem = 213 * 1024;
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("cluster size %2d: max active clusters %d (= %d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaGetLastError();
    cudaFuncSetAttribute(lat, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* out; unsigned* ctr;
    cudaMalloc(&out, 4 * 160 * sizeof(long long));
    cudaMalloc(&ctr, 64 * sizeof(unsigned));
    for (int cs : {8, 4}) {
        for (int grid : {128}) {
            cudaMemset(ctr, 0, 64 * sizeof(unsigned));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[2];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
            cfg.attrs = at; cfg.numAttrs = 2;
            cudaError_t e = cudaLaunchKernelEx(&cfg, lat, out, ctr);
            cudaError_t e2 = cudaDeviceSynchronize();
            printf("lat cs=%d grid=%d: launch %s sync %s\n", cs, grid, cudaGetErrorString(e), cudaGetErrorString(e2));
            long long h[4 * 160];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("  block 0: cluster.sync %lld cyc, DSMEM dependent load %lld cyc, grid barrier(128 atomics) %lld cyc, hierarchical %lld cyc\n", h[0], h[1], h[2], h[3]);
            printf("  block 77: %lld %lld %lld %lld\n", h[77 * 4], h[77 * 4 + 1], h[77 * 4 + 2], h[77 * 4 + 3]);
        }
    }
    return 0;
}
