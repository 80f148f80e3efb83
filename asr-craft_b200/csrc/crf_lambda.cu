// Device-side derivation of everything the kernels read from lambda, and the trainer's per-minibatch update, so that lambda can stay
// resident in HBM between minibatches (crfgpu_sgd_update) and crfgpu_set_lambda costs one copy instead of ~7 ms of host loops:
//   state weights / biases as floats, E = exp(M - Mmax) and its transpose (CRF_StdFeatureMap::computeTransMatrixValue with bias-only
//   transitions, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:94-110), the bf16 hi/lo UMMA weight tiles of the TMA-fed score GEMM, and the
//   decoder's tables in the reference's arithmetic (double score, negate, narrow to float:
//   CRF/src/decoders/CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:286,315,331,458).
// Update rule: CRF_SGTrainer::sgtrainMinibatch, CRF/src/trainers/CRF_SGTrainer.cpp:299-325 (plain SGD, AdaGrad, the `grad -= grad/gvar`
// quirk, lambdaAcc / lambdaSqrAcc for the averaged model).
#include <cfloat>

#include <cuda_bf16.h>

#include "crf_kernels.cuh"

namespace crfgpu {

namespace {

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
	unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
	unsigned long long old = *a;
	while (__longlong_as_double((long long)old) < v) {
		const unsigned long long assumed = old;
		old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
		if (old == assumed) break;
	}
}

__global__ void __launch_bounds__(256) lambda_tmax_kernel(LambdaTablesParams p) {
	__shared__ double sm[8];
	double m = -DBL_MAX;
	const uint64_t n = (uint64_t)p.Le * p.Le;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t ti = p.tidx[i];
		if (ti != 0xffffffffu) m = fmax(m, p.use_trans_bias ? __dmul_rn(p.lam[ti], p.trans_bias_val) : 0.0);
	}
	for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; w++) m = fmax(m, sm[w]);
		if (m > -DBL_MAX) atomic_max_double(p.tmax, m);
	}
}

__global__ void __launch_bounds__(256) lambda_tables_kernel(LambdaTablesParams p) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, tid0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	double tmax = *p.tmax;
	if (tmax == -DBL_MAX) tmax = 0.0;
	if (p.Ws) {
		const uint64_t nw = (uint64_t)p.Le * p.nSf;
		for (uint64_t i = tid0; i < nw; i += stride) {
			const uint32_t cl = (uint32_t)(i / p.nSf), f = (uint32_t)(i % p.nSf);
			p.Ws[i] = (float)p.lam[p.sidx[cl] + f];
		}
		for (uint64_t cl = tid0; cl < p.Le; cl += stride) p.bias[cl] = p.use_state_bias ? (float)__dmul_rn(p.lam[p.sidx[cl] + p.nSf], p.state_bias_val) : 0.0f;
		const uint64_t n = (uint64_t)p.Le * p.Le;
		for (uint64_t i = tid0; i < n; i += stride) {
			const uint32_t ti = p.tidx[i];
			if (ti == 0xffffffffu) continue;               // illegal pair: E stays 0 (the tables are zeroed before the launch)
			const uint32_t q = (uint32_t)(i / p.Le), cl = (uint32_t)(i % p.Le);
			const float e = (float)exp((p.use_trans_bias ? __dmul_rn(p.lam[ti], p.trans_bias_val) : 0.0) - tmax);
			p.E[(uint64_t)q * p.Lpe + cl] = e; p.ET[(uint64_t)cl * p.Lpe + q] = e;
		}
	}
	if (p.Wt) {
		// tile (d, jt, c): [hi 4 KB | lo 4 KB], element (r, kk) at (r/8)*512 + (kk/8)*128 + (r%8)*16 + (kk%8)*2   (split_weight_tiles)
		const uint32_t ntile = (p.wt_P + 63) / 64;
		const uint64_t n = (uint64_t)p.wt_D * ntile * p.wt_chunks * 64 * 32;
		for (uint64_t i = tid0; i < n; i += stride) {
			const uint32_t kk = (uint32_t)(i & 31), r = (uint32_t)((i >> 5) & 63);
			const uint64_t tile = i >> 11;
			const uint32_t c = (uint32_t)(tile % p.wt_chunks), jt = (uint32_t)((tile / p.wt_chunks) % ntile), d = (uint32_t)(tile / ((uint64_t)p.wt_chunks * ntile));
			const uint32_t y = jt * 64 + r;
			uint32_t k = c * 32 + kk;
			if (p.wt_virt) {
				// chunk order of the virtual windows -> feature index inside the window (0xffffffff: padding)
				if (c < 5 * p.wt_cpb) { const uint32_t b = c / p.wt_cpb, f = (c - b * p.wt_cpb) * 32 + kk; k = f < p.wt_F ? b * p.wt_F + f : 0xffffffffu; }
				else { const uint32_t a = (c - 5 * p.wt_cpb) * 32 + kk; k = a < 3 * p.wt_F ? 5 * p.wt_F + a : 0xffffffffu; }
			}
			float w = 0.0f;
			if (y < p.wt_P && k < p.nSf) w = (float)p.lam[p.sidx[d * p.wt_P + y] + k];
			const __nv_bfloat16 hi = __float2bfloat16_rn(w), lo = __float2bfloat16_rn(w - __bfloat162float(hi));
			unsigned char* t = p.Wt + tile * 8192 + (r / 8) * 512 + (kk / 8) * 128 + (r % 8) * 16 + (kk % 8) * 2;
			*reinterpret_cast<__nv_bfloat16*>(t) = hi; *reinterpret_cast<__nv_bfloat16*>(t + 4096) = lo;
		}
	}
	if (p.bias_dy) {
		// per-(duration, label) bias of the virtual windows: the state bias + the weight of the one-hot duration feature of duration d
		const uint64_t n = (uint64_t)p.wt_Dd * p.wt_P;
		for (uint64_t i = tid0; i < n; i += stride) {
			const uint32_t d = (uint32_t)(i / p.wt_P), y = (uint32_t)(i % p.wt_P);
			const uint32_t sb = p.sidx[(p.wt_D == 1 ? 0 : d * p.wt_P) + y];
			const double b = p.use_state_bias ? __dmul_rn(p.lam[sb + p.nSf], p.state_bias_val) : 0.0;
			p.bias_dy[i] = (float)(b + p.lam[sb + 8 * p.wt_F + d]);
		}
	}
	if (p.Wtr) {
		const uint64_t L = p.L0, nw = L * L * p.nTf;
		for (uint64_t i = tid0; i < nw; i += stride) {
			const uint64_t pair = i / p.nTf; const uint32_t f = (uint32_t)(i % p.nTf);
			const uint32_t b = p.tidx0[pair];
			p.Wtr[i] = b != 0xffffffffu ? (float)p.lam[b + f] : 0.0f;
		}
		for (uint64_t pair = tid0; pair < L * L; pair += stride)
			// illegal pairs of an N-state map: score -inf, so that exp(M - max) is an exact 0 in the recursions
			p.tbias[pair] = p.tidx0[pair] == 0xffffffffu ? -INFINITY : (p.use_trans_bias ? (float)__dmul_rn(p.lam[p.tidx0[pair] + p.nTf], p.trans_bias_val) : 0.0f);
	}
	if (p.WdT) {
		const uint64_t nw = (uint64_t)(p.nTf + 1) * p.vtE;
		for (uint64_t i = tid0; i < nw; i += stride) {
			const uint32_t f = (uint32_t)(i / p.vtE), e = (uint32_t)(i % p.vtE);
			const uint32_t b = p.vt_base[e];
			p.WdT[i] = (b != 0xffffffffu && (f < p.nTf || p.use_trans_bias)) ? p.lam[b + f] : 0.0;
		}
	}
	if (p.Wd) {
		// decoder tables over the MODEL's labels (L0 labels, NS sub-states, P0 phones)
		const uint32_t L = p.L0, NS = p.NS, P = p.P0;
		const uint64_t nw = (uint64_t)(p.nSf + 1) * L;
		for (uint64_t i = tid0; i < nw; i += stride) {
			const uint32_t f = (uint32_t)(i / L), cl = (uint32_t)(i % L);
			p.Wd[i] = (f < p.nSf || p.use_state_bias) ? p.lam[p.sidx0[cl] + f] : 0.0;
		}
		auto Mv = [&](uint32_t a, uint32_t b) -> double {
			const uint32_t ti = p.tidx0[(uint64_t)a * L + b];
			return (ti != 0xffffffffu && p.use_trans_bias) ? __dmul_rn(p.lam[ti], p.trans_bias_val) : 0.0;
		};
		for (uint64_t i = tid0; i < (uint64_t)P * P; i += stride) {
			const uint32_t pp = (uint32_t)(i / P), q = (uint32_t)(i % P);
			p.crossT[i] = (float)(-1 * Mv(pp * NS + NS - 1, q * NS));
		}
		for (uint64_t l = tid0; l < L; l += stride) {
			p.negDiag[l] = (float)(-1 * Mv((uint32_t)l, (uint32_t)l));
			p.negOff[l] = (l % NS != 0) ? (float)(-1 * Mv((uint32_t)l - 1, (uint32_t)l)) : 0.0f;
		}
	}
}

__global__ void __launch_bounds__(256) sgd_update_kernel(SgdParams p) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < p.len; i += (uint64_t)gridDim.x * blockDim.x) {
		double g = p.grad[i] / p.n_active;                                  // CRF_Minibatch_GradAccumulator.cpp:306-308
		if (p.use_gvar) g = __dsub_rn(g, __dmul_rn(g, p.inv_square_var));   // CRF_SGTrainer.cpp:300-303 (sic: grad, not lambda)
		double l = p.lambda[i];
		if (p.use_adagrad) {
			const double a = __dadd_rn(p.grad_sqr_acc[i], __dmul_rn(g, g));
			p.grad_sqr_acc[i] = a;
			l = __dadd_rn(l, __dmul_rn(p.eta / (sqrt(a) + p.eps), g));
		} else l = __dadd_rn(l, __dmul_rn(p.lr, g));
		p.lambda[i] = l;
		if (p.lambda_acc) p.lambda_acc[i] = __dadd_rn(p.lambda_acc[i], l);
		if (p.lambda_sqr_acc) p.lambda_sqr_acc[i] = __dadd_rn(p.lambda_sqr_acc[i], __dmul_rn(l, l));
	}
}

}  // namespace

cudaError_t launch_lambda_tables(const LambdaTablesParams& p, cudaStream_t s) {
	const double init = -DBL_MAX;
	cudaError_t e = cudaMemcpyAsync(p.tmax, &init, sizeof(double), cudaMemcpyHostToDevice, s);
	if (e != cudaSuccess) return e;
	if (p.Ws) lambda_tmax_kernel<<<148 * 4, 256, 0, s>>>(p);
	lambda_tables_kernel<<<148 * 8, 256, 0, s>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_sgd_update(const SgdParams& p, cudaStream_t s) {
	if (!p.len) return cudaSuccess;
	sgd_update_kernel<<<148 * 4, 256, 0, s>>>(p);
	return cudaGetLastError();
}

}  // namespace crfgpu
