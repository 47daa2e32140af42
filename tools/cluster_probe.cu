// Developer probe: where do the two CTAs of a 2-CTA cluster land? Prints, per SM, how many rank-0 and rank-1 CTAs it ran.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__global__ void __cluster_dims__(2, 1, 1) probe(int *smid, int *rank, int spin)
{
    unsigned s, r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    if (threadIdx.x == 0)
    {
        smid[blockIdx.x] = (int)s;
        rank[blockIdx.x] = (int)r;
    }
    long long t0 = clock64();
    while (clock64() - t0 < spin) // keep the CTAs resident so that the whole grid is placed at once
        ;
    cg::this_cluster().sync();
}
int main()
{
    for (int regs_heavy = 0; regs_heavy < 2; ++regs_heavy)
        for (int grid : {148, 296, 592})
        {
            int *smid, *rank;
            cudaMalloc(&smid, grid * 4);
            cudaMalloc(&rank, grid * 4);
            probe<<<grid, 256, regs_heavy ? 100 * 1024 : 0>>>(smid, rank, 2000000);
            cudaError_t e = cudaDeviceSynchronize();
            int hs[1024], hr[1024];
            cudaMemcpy(hs, smid, grid * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hr, rank, grid * 4, cudaMemcpyDeviceToHost);
            int cnt[256][2] = {};
            for (int i = 0; i < grid; ++i)
                cnt[hs[i]][hr[i]]++;
            int both = 0, only0 = 0, only1 = 0, pairs_adjacent = 0;
            for (int s = 0; s < 256; ++s)
            {
                if (cnt[s][0] && cnt[s][1]) both++;
                else if (cnt[s][0]) only0++;
                else if (cnt[s][1]) only1++;
            }
            for (int i = 0; i + 1 < grid; i += 2)
                if ((hs[i] ^ hs[i + 1]) == 1) pairs_adjacent++;
            printf("grid %d smem %d KB (%s): SMs with both ranks %d, only rank0 %d, only rank1 %d; clusters on an (even, odd) SM pair: %d of %d\n", grid,
                   regs_heavy ? 100 : 0, cudaGetErrorString(e), both, only0, only1, pairs_adjacent, grid / 2);
            if (grid == 296 && !regs_heavy)
            {
                printf("first clusters (smid rank): ");
                for (int i = 0; i < 16; ++i) printf("(%d %d) ", hs[i], hr[i]);
                printf("\n");
            }
        }
    return 0;
}
