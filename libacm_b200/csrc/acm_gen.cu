/*
 * acm_gen.cu -- on-GPU corpus generation.
 *
 * markokr/libacm ships no encoder and no sample files; every stream this repository decodes is
 * synthetic (acmgen.c).  BASELINE configs[3] asks for 1 000 000 streams per job: generating
 * them on the host and copying them over PCIe costs minutes and tens of GB.  Here the very
 * same generator (acmgen_core.h, compiled as device code) runs one thread per stream and
 * writes the images straight into HBM: pass 1 sizes every image, the host lays them out
 * back-to-back on 16-byte boundaries (what acmgen_write_many does), pass 2 writes them.
 * The images are byte-identical to the host generator's (tests/test_gpu_gen.py).
 *
 * This is test / benchmark infrastructure on the input side of the decode path; it is not part
 * of the reference's API surface.
 */
#include <cuda_runtime.h>

#include <vector>

#include "acm_gpu.h"
#include "acm_host.h"
#include "libacm.h"
#include "acmgen_core.h"

namespace {

constexpr int GEN_THREADS = 64;
constexpr uint32_t GEN_MAX_LEVEL = 10; /* per-thread selector scratch: 1 << level bytes */

__global__ void __launch_bounds__(GEN_THREADS) acmgen_size_kernel(const acmgen_params *params, uint64_t n, uint32_t *lens)
{
	const uint64_t i = (uint64_t)blockIdx.x * GEN_THREADS + threadIdx.x;
	uint8_t inds[1u << GEN_MAX_LEVEL];
	if (i >= n)
		return;
	const acmgen_params p = params[i];
	lens[i] = (uint32_t)acmgen_write_core(&p, nullptr, 0, inds, 1);
}

__global__ void __launch_bounds__(GEN_THREADS)
acmgen_write_kernel(const acmgen_params *params, uint64_t n, uint8_t *blob, const uint64_t *offs, uint32_t *lens)
{
	const uint64_t i = (uint64_t)blockIdx.x * GEN_THREADS + threadIdx.x;
	uint8_t inds[1u << GEN_MAX_LEVEL];
	if (i >= n)
		return;
	const acmgen_params p = params[i];
	/* lens[i] is the size pass 1 found; 0 afterwards flags a mismatch (cannot happen: same code) */
	const size_t got = acmgen_write_core(&p, blob + offs[i], lens[i], inds, 0);
	if (got != lens[i])
		lens[i] = 0;
}

#define GEN_CUDA(call)                                                                                  \
	do {                                                                                            \
		cudaError_t e_ = (call);                                                                \
		if (e_ != cudaSuccess) {                                                                \
			acm_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			rc = ACM_ERR_OTHER;                                                             \
			goto done;                                                                      \
		}                                                                                       \
	} while (0)

} // namespace

extern "C" int acm_gpu_generate(const void *params_v, uint64_t n, void *d_blob, uint64_t cap, uint64_t *offs,
				uint32_t *lens, uint64_t *used, int device)
{
	const acmgen_params *params = static_cast<const acmgen_params *>(params_v);
	int rc = ACM_OK;
	acmgen_params *d_params = nullptr;
	uint32_t *d_lens = nullptr;
	uint64_t *d_offs = nullptr;
	uint64_t at = 0;
	const unsigned grid = (unsigned)((n + GEN_THREADS - 1) / GEN_THREADS);
	if (!params || !offs || !lens || !used || n == 0) {
		acm_set_error("acm_gpu_generate: missing argument");
		return ACM_ERR_OTHER;
	}
	for (uint64_t i = 0; i < n; i++) {
		if (params[i].level > GEN_MAX_LEVEL || params[i].rows == 0 || params[i].rows > 4095) {
			acm_set_error("acm_gpu_generate: stream %llu: level %u / rows %u not supported on the device (level <= %u)",
				      (unsigned long long)i, params[i].level, params[i].rows, GEN_MAX_LEVEL);
			return ACM_ERR_OTHER;
		}
	}
	if (device >= 0) {
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess) {
			acm_set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
			return ACM_ERR_OTHER;
		}
	}
	GEN_CUDA(cudaMalloc(&d_params, n * sizeof(acmgen_params)));
	GEN_CUDA(cudaMalloc(&d_lens, n * sizeof(uint32_t)));
	GEN_CUDA(cudaMalloc(&d_offs, n * sizeof(uint64_t)));
	GEN_CUDA(cudaMemcpy(d_params, params, n * sizeof(acmgen_params), cudaMemcpyHostToDevice));
	acmgen_size_kernel<<<grid, GEN_THREADS>>>(d_params, n, d_lens);
	GEN_CUDA(cudaGetLastError());
	GEN_CUDA(cudaMemcpy(lens, d_lens, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	for (uint64_t i = 0; i < n; i++) {
		at = (at + 15) & ~(uint64_t)15;
		offs[i] = at;
		at += lens[i];
	}
	*used = at;
	if (d_blob == nullptr)
		goto done; /* sizing call */
	if (at > cap) {
		acm_set_error("acm_gpu_generate: the corpus needs %llu bytes, the blob has %llu", (unsigned long long)at,
			      (unsigned long long)cap);
		rc = ACM_ERR_OTHER;
		goto done;
	}
	GEN_CUDA(cudaMemcpy(d_offs, offs, n * sizeof(uint64_t), cudaMemcpyHostToDevice));
	/* the gaps between images read as zero, like the host generator's calloc'ed blob */
	GEN_CUDA(cudaMemset(d_blob, 0, (size_t)at));
	acmgen_write_kernel<<<grid, GEN_THREADS>>>(d_params, n, static_cast<uint8_t *>(d_blob), d_offs, d_lens);
	GEN_CUDA(cudaGetLastError());
	{
		std::vector<uint32_t> check(n);
		GEN_CUDA(cudaMemcpy(check.data(), d_lens, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
		for (uint64_t i = 0; i < n; i++) {
			if (check[i] != lens[i]) {
				acm_set_error("acm_gpu_generate: stream %llu changed size between the passes",
					      (unsigned long long)i);
				rc = ACM_ERR_OTHER;
				break;
			}
		}
	}
done:
	cudaFree(d_params);
	cudaFree(d_lens);
	cudaFree(d_offs);
	return rc;
}
