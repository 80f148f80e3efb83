// dsmem_probe: cost of the per-frame all-gather inside a thread-block cluster (the exchange step of the
// cluster-resident lattice kernels).  Every CTA of an 8-CTA cluster owns a slice of SLICE bytes and
// delivers it to all 8 CTAs each iteration; receivers wait on a local mbarrier.  Variants:
//   0: st.shared::cluster.v4 by all threads + one remote mbarrier.arrive.release.cluster per destination
//   1: cp.async.bulk.shared::cluster.shared::cta (one bulk copy per destination, complete_tx on the remote mbarrier)
//   2: st.global slice + barrier.cluster + every CTA reads all slices back from L2 (the round-1 scheme)
// Development tool:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o dsmem_probe dsmem_probe.cu
#include <cstdio>
#include <cstdlib>

#include "../asr-craft_b200/csrc/tc05.cuh"

using namespace tc05;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int CS = 8;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
	uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) allgather_kernel(int slice_bytes, int iters, long long* cycles, float* gbuf, float* sink) {
	extern __shared__ __align__(128) unsigned char smem[];
	// [2 buffers][CS slices][slice_bytes] receive area, then my staging slice
	unsigned char* recv = smem;
	unsigned char* stage = smem + 2 * CS * slice_bytes;
	__shared__ uint64_t bar[2];
	const uint32_t rank = ctarank(), tid = threadIdx.x;
	const uint32_t cl = blockIdx.x / CS;
	if (tid == 0) { mbar_init(&bar[0], VARIANT == 0 ? CS : 1); mbar_init(&bar[1], VARIANT == 0 ? CS : 1); fence_mbar_init(); }
	for (int i = tid; i < slice_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(stage)[i] = (float)(rank * 1000 + i);
	__syncthreads();
	cluster_sync_all();
	float acc = 0.0f;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		const int b = it & 1;
		unsigned char* rb = recv + (size_t)b * CS * slice_bytes;
		if (VARIANT == 0) {
			const uint32_t dst_local = smem_u32(rb + rank * slice_bytes);
			for (int i = tid; i < slice_bytes / 16; i += blockDim.x) {
				const uint4 v = reinterpret_cast<const uint4*>(stage)[i];
#pragma unroll
				for (int r = 0; r < CS; r++) {
					const uint32_t ra = mapa(dst_local + i * 16, r);
					asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
				}
			}
			__syncthreads();
			if (tid < CS) {
				const uint32_t rbar = mapa(smem_u32(&bar[b]), tid);
				asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
			}
			mbar_wait_cluster(&bar[b], (it >> 1) & 1);
		} else if (VARIANT == 1) {
			if (tid == 0) mbar_arrive_expect_tx(&bar[b], CS * slice_bytes);
			fence_proxy_async_smem();
			__syncthreads();
			if (tid < CS) {
				const uint32_t dst = mapa(smem_u32(rb + rank * slice_bytes), tid);
				const uint32_t rbar = mapa(smem_u32(&bar[b]), tid);
				asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				             ::"r"(dst), "r"(smem_u32(stage)), "r"(slice_bytes), "r"(rbar) : "memory");
			}
			mbar_wait_cluster(&bar[b], (it >> 1) & 1);
		} else {
			float* g = gbuf + ((size_t)cl * 2 + b) * CS * (slice_bytes / 4);
			for (int i = tid; i < slice_bytes / 16; i += blockDim.x)
				reinterpret_cast<uint4*>(g + (size_t)rank * (slice_bytes / 4))[i] = reinterpret_cast<const uint4*>(stage)[i];
			cluster_sync_all();
			for (int i = tid; i < CS * slice_bytes / 16; i += blockDim.x)
				reinterpret_cast<uint4*>(rb)[i] = __ldcg(reinterpret_cast<const uint4*>(g) + i);
			__syncthreads();
		}
		// consume something so the data really has to be there
		acc += reinterpret_cast<const float*>(rb)[(tid * 37 + it) % (CS * slice_bytes / 4)];
		// next iteration's payload depends on what arrived (models the recursion's dependency)
		if (tid == 0) reinterpret_cast<float*>(stage)[0] = acc * 1e-30f + (float)(rank * 1000);
		__syncthreads();
	}
	long long t1 = clock64();
	cluster_sync_all();
	if (tid == 0 && rank == 0) cycles[cl] = t1 - t0;
	if (acc == 123.456f) sink[0] = acc;
	// verify last buffer: slice r element i == r*1000+i (i>0)
	if (blockIdx.x == 0 && tid == 1) {
		const float* rb = reinterpret_cast<const float*>(recv + (size_t)((iters - 1) & 1) * CS * slice_bytes);
		int bad = 0;
		for (int r = 0; r < CS; r++) for (int i = 1; i < slice_bytes / 4; i++) if (rb[r * (slice_bytes / 4) + i] != (float)(r * 1000 + i)) bad++;
		sink[1] = (float)bad;
	}
}

template <int VARIANT>
static void run(const char* name, int slice_bytes, int n_clusters) {
	const int iters = 400;
	long long* cyc; float *gbuf, *sink;
	CK(cudaMalloc(&cyc, sizeof(long long) * n_clusters)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
	CK(cudaMalloc(&gbuf, (size_t)n_clusters * 2 * CS * slice_bytes));
	const size_t smem = (size_t)(2 * CS + 1) * slice_bytes;
	auto kern = allgather_kernel<VARIANT>;
	CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(n_clusters * CS); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	CK(cudaLaunchKernelEx(&cfg, kern, slice_bytes, iters, cyc, gbuf, sink));
	CK(cudaDeviceSynchronize());
	long long h[64]; float hs[2];
	CK(cudaMemcpy(h, cyc, sizeof(long long) * n_clusters, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hs, sink, 8, cudaMemcpyDeviceToHost));
	long long mx = 0; for (int i = 0; i < n_clusters; i++) mx = h[i] > mx ? h[i] : mx;
	printf("%-28s slice %6d B (gather %7d B/CTA)  clusters %2d  %8.0f cycles/iter   bad=%d\n", name, slice_bytes, CS * slice_bytes, n_clusters,
	       (double)mx / iters, (int)hs[1]);
	cudaFree(cyc); cudaFree(gbuf); cudaFree(sink);
}

int main() {
	for (int ncl : {1, 15}) {
		for (int sb : {2560, 5120, 10240}) {
			run<0>("st.shared::cluster", sb, ncl);
			run<1>("cp.async.bulk smem->dsmem", sb, ncl);
			run<2>("global + barrier.cluster", sb, ncl);
		}
	}
	return 0;
}
