// tc_probe: hardware check of the tcgen05 building blocks in csrc/tc05.cuh (descriptor encodings, the
// no-swizzle canonical layouts in both majors, TMEM lane mapping, split-bf16 accuracy).  Development tool,
// not part of the product library:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../asr-craft_b200/csrc/tc05.cuh"

using namespace tc05;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// D[128 x N] = A[128 x K] * B[N x K]^T, A/B given row-major fp32 (A[m][k], B[n][k]); ROWS_VALID < 128 leaves the
// remaining A rows unwritten (garbage lanes) to check that they do not disturb the valid ones.
template <int N, int K, int AROWS, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, int rows_valid) {
	extern __shared__ __align__(128) unsigned char smem[];
	constexpr int KG = K / 8;
	constexpr uint32_t A_BYTES = AROWS * K * 2, B_BYTES = N * K * 2;
	unsigned char* a_hi = smem; unsigned char* a_lo = a_hi + A_BYTES;
	unsigned char* b_hi = a_lo + A_BYTES; unsigned char* b_lo = b_hi + B_BYTES;
	__shared__ uint64_t bar;
	__shared__ uint32_t tmem_base;
	const int tid = threadIdx.x, warp = tid >> 5;
	// K-major: k-groups contiguous (LBO=128), row groups at KG*128.  MN-major: row groups contiguous (SBO=128), k-groups at (R/8)*128.
	constexpr uint32_t A_LBO = A_MN ? (AROWS / 8) * 128 : 128, A_SBO = A_MN ? 128 : KG * 128;
	constexpr uint32_t B_LBO = B_MN ? (N / 8) * 128 : 128, B_SBO = B_MN ? 128 : KG * 128;
	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	if (warp == 0) tmem_alloc(&tmem_base, N < 32 ? 32 : N);
	// fill: one (row, k-group) 16-byte vector per iteration for K-major; one (row-group, k) for MN-major
	for (int i = tid; i < AROWS * KG; i += 128) {
		float x[8]; uint4 h, l;
		if (!A_MN) {
			const int r = i / KG, kg = i % KG;
			if (r >= rows_valid) continue;
			for (int j = 0; j < 8; j++) x[j] = A[r * K + kg * 8 + j];
			split8(x, h, l);
			const uint32_t o = (r / 8) * A_SBO + kg * A_LBO + (r % 8) * 16;
			*reinterpret_cast<uint4*>(a_hi + o) = h; *reinterpret_cast<uint4*>(a_lo + o) = l;
		} else {
			const int rg = i % (AROWS / 8), k = i / (AROWS / 8);
			if (rg * 8 >= rows_valid) continue;
			for (int j = 0; j < 8; j++) x[j] = (rg * 8 + j < rows_valid) ? A[(rg * 8 + j) * K + k] : 0.0f;
			split8(x, h, l);
			const uint32_t o = rg * A_SBO + (k / 8) * A_LBO + (k % 8) * 16;
			*reinterpret_cast<uint4*>(a_hi + o) = h; *reinterpret_cast<uint4*>(a_lo + o) = l;
		}
	}
	for (int i = tid; i < N * KG; i += 128) {
		float x[8]; uint4 h, l;
		if (!B_MN) {
			const int r = i / KG, kg = i % KG;
			for (int j = 0; j < 8; j++) x[j] = B[r * K + kg * 8 + j];
			split8(x, h, l);
			const uint32_t o = (r / 8) * B_SBO + kg * B_LBO + (r % 8) * 16;
			*reinterpret_cast<uint4*>(b_hi + o) = h; *reinterpret_cast<uint4*>(b_lo + o) = l;
		} else {
			const int rg = i % (N / 8), k = i / (N / 8);
			for (int j = 0; j < 8; j++) x[j] = B[(rg * 8 + j) * K + k];
			split8(x, h, l);
			const uint32_t o = rg * B_SBO + (k / 8) * B_LBO + (k % 8) * 16;
			*reinterpret_cast<uint4*>(b_hi + o) = h; *reinterpret_cast<uint4*>(b_lo + o) = l;
		}
	}
	fence_proxy_async_smem();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tb = tmem_base;
	if (tid == 0) {
		constexpr uint32_t idesc = idesc_bf16_f32(128, N, A_MN, B_MN);
		for (int ks = 0; ks < K / 16; ks++) {
			const uint32_t ao = ks * 2 * A_LBO, bo = ks * 2 * B_LBO;
			const uint64_t dah = smem_desc(smem_u32(a_hi) + ao, A_LBO, A_SBO), dal = smem_desc(smem_u32(a_lo) + ao, A_LBO, A_SBO);
			const uint64_t dbh = smem_desc(smem_u32(b_hi) + bo, B_LBO, B_SBO), dbl = smem_desc(smem_u32(b_lo) + bo, B_LBO, B_SBO);
			mma_ss(tb, dah, dbh, idesc, ks > 0);
			mma_ss(tb, dal, dbh, idesc, true);
			mma_ss(tb, dah, dbl, idesc, true);
		}
		mma_commit(&bar);
	}
	mbar_wait(&bar, 0);
	tc_fence_after();
	for (int c0 = 0; c0 < N; c0 += 8) {
		float v[8];
		tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
		tmem_ld_wait();
		for (int j = 0; j < 8; j++) D[tid * N + c0 + j] = v[j];
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) tmem_dealloc(tb, N < 32 ? 32 : N);
}

// TS mode: A_hi lives in TMEM (lane = row, column j of a k-step = bf16x2 (k=2j, 2j+1)), A_lo and B hi/lo in shared memory.
template <int N, int K>
__global__ void __launch_bounds__(128) probe_ts_kernel(const float* A, const float* B, float* D, int rows_valid) {
	extern __shared__ __align__(128) unsigned char smem[];
	constexpr int KG = K / 8;
	constexpr uint32_t A_BYTES = 128 * K * 2, B_BYTES = N * K * 2;
	unsigned char* a_lo = smem; unsigned char* b_hi = a_lo + A_BYTES; unsigned char* b_lo = b_hi + B_BYTES;
	__shared__ uint64_t bar;
	__shared__ uint32_t tmem_base;
	const int tid = threadIdx.x, warp = tid >> 5;
	constexpr uint32_t LBO = 128, SBO = KG * 128;
	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	if (warp == 0) tmem_alloc(&tmem_base, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tb = tmem_base;
	const uint32_t d_col = 0, a_col = 32;            // accumulator at column 0, A_hi from column 32
	// my row of A: hi -> TMEM, lo -> smem
	for (int kg = 0; kg < KG; kg++) {
		float x[8]; uint4 h, l;
		for (int j = 0; j < 8; j++) x[j] = tid < rows_valid ? A[tid * K + kg * 8 + j] : 0.0f;
		split8(x, h, l);
		*reinterpret_cast<uint4*>(a_lo + (tid / 8) * SBO + kg * LBO + (tid % 8) * 16) = l;
		// 8 k values = 4 packed columns; write two k-groups (8 columns) at a time
		static uint32_t dummy;
		(void)dummy;
		uint32_t r[8] = {h.x, h.y, h.z, h.w, 0, 0, 0, 0};
		if (kg + 1 < KG) {
			float y[8]; uint4 h2, l2;
			for (int j = 0; j < 8; j++) y[j] = tid < rows_valid ? A[tid * K + (kg + 1) * 8 + j] : 0.0f;
			split8(y, h2, l2);
			*reinterpret_cast<uint4*>(a_lo + (tid / 8) * SBO + (kg + 1) * LBO + (tid % 8) * 16) = l2;
			r[4] = h2.x; r[5] = h2.y; r[6] = h2.z; r[7] = h2.w;
		}
		tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + a_col + kg * 4, r);
		kg++;
	}
	tmem_st_wait();
	for (int i = tid; i < N * KG; i += 128) {
		float x[8]; uint4 h, l;
		const int r = i / KG, kg = i % KG;
		for (int j = 0; j < 8; j++) x[j] = B[r * K + kg * 8 + j];
		split8(x, h, l);
		const uint32_t o = (r / 8) * SBO + kg * LBO + (r % 8) * 16;
		*reinterpret_cast<uint4*>(b_hi + o) = h; *reinterpret_cast<uint4*>(b_lo + o) = l;
	}
	fence_proxy_async_smem();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	if (tid == 0) {
		constexpr uint32_t idesc = idesc_bf16_f32(128, N, false, false);
		for (int ks = 0; ks < K / 16; ks++) {
			const uint64_t dal = smem_desc(smem_u32(a_lo) + ks * 256, LBO, SBO);
			const uint64_t dbh = smem_desc(smem_u32(b_hi) + ks * 256, LBO, SBO), dbl = smem_desc(smem_u32(b_lo) + ks * 256, LBO, SBO);
			mma_ts(tb + d_col, tb + a_col + ks * 8, dbh, idesc, ks > 0);
			mma_ts(tb + d_col, tb + a_col + ks * 8, dbl, idesc, true);
			mma_ss(tb + d_col, dal, dbh, idesc, true);
		}
		mma_commit(&bar);
	}
	mbar_wait(&bar, 0);
	tc_fence_after();
	for (int c0 = 0; c0 < N; c0 += 8) {
		float v[8];
		tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + d_col + c0, v);
		tmem_ld_wait();
		for (int j = 0; j < 8; j++) D[tid * N + c0 + j] = v[j];
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int K>
static bool run_ts_case(const char* name, int rows_valid) {
	std::vector<float> A(128 * K), B(N * K), D(128 * N, -1.0f);
	srand(4321 + N + K);
	for (auto& v : A) v = expf(-3.0f * (float)rand() / RAND_MAX);
	for (auto& v : B) v = (float)rand() / RAND_MAX;
	float *dA, *dB, *dD;
	CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
	CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
	const size_t smem = 128 * K * 2 + 2 * (N * K * 2) + 1024;
	auto kern = probe_ts_kernel<N, K>;
	CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	kern<<<1, 128, smem>>>(dA, dB, dD, rows_valid);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
	double worst = 0.0;
	for (int m = 0; m < rows_valid; m++)
		for (int n = 0; n < N; n++) {
			double ref = 0.0;
			for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * (double)B[n * K + k];
			worst = fmax(worst, fabs(D[m * N + n] - ref) / fabs(ref));
		}
	const bool ok = worst < 2e-5;
	printf("%-44s rows=%3d  worst rel err %.3e  %s\n", name, rows_valid, worst, ok ? "OK" : "FAIL");
	cudaFree(dA); cudaFree(dB); cudaFree(dD);
	return ok;
}

template <int N, int K, int AROWS, bool A_MN, bool B_MN>
static bool run_case(const char* name, int rows_valid) {
	std::vector<float> A(128 * K), B(N * K), D(128 * N, -1.0f);
	srand(1234 + N + K);
	for (auto& v : A) v = (float)rand() / RAND_MAX;                  // probabilities in [0,1]
	for (auto& v : B) v = expf(-3.0f * (float)rand() / RAND_MAX);    // exp(M - Mmax) in (0.05, 1]
	float *dA, *dB, *dD;
	CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
	CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemset(dD, 0xff, D.size() * 4));
	const size_t smem = 2 * (AROWS * K * 2) + 2 * (N * K * 2) + 16384;   // slack: garbage row groups of the 80-row cases read past the tiles
	auto kern = probe_kernel<N, K, AROWS, A_MN, B_MN>;
	CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	kern<<<1, 128, smem>>>(dA, dB, dD, rows_valid);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
	double worst = 0.0;
	for (int m = 0; m < rows_valid; m++)
		for (int n = 0; n < N; n++) {
			double ref = 0.0;
			for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * (double)B[n * K + k];
			worst = fmax(worst, fabs(D[m * N + n] - ref) / fabs(ref));
		}
	const bool ok = worst < 2e-5;
	printf("%-44s rows=%3d  worst rel err %.3e  %s\n", name, rows_valid, worst, ok ? "OK" : "FAIL");
	cudaFree(dA); cudaFree(dB); cudaFree(dD);
	return ok;
}

int main() {
	bool ok = true;
	ok &= run_case<64, 64, 128, false, false>("M128 N64 K64  A K-major  B K-major", 128);
	ok &= run_case<64, 64, 128, true, true>("M128 N64 K64  A MN-major B MN-major", 128);
	ok &= run_case<64, 64, 128, true, false>("M128 N64 K64  A MN-major B K-major", 128);
	ok &= run_case<16, 128, 128, false, false>("M128 N16 K128 A K-major  B K-major", 128);
	ok &= run_case<16, 128, 80, false, false>("M128 N16 K128 A K-major  B K-major (80 rows)", 80);
	ok &= run_case<32, 448, 80, false, false>("M128 N32 K448 A K-major  B K-major (80 rows)", 80);
	ok &= run_ts_case<16, 128>("TS: A_hi in TMEM, M128 N16 K128", 128);
	ok &= run_ts_case<16, 640>("TS: A_hi in TMEM, M128 N16 K640 (80 rows)", 80);
	printf(ok ? "tc_probe: ALL OK\n" : "tc_probe: FAILURES\n");
	return ok ? 0 : 1;
}
