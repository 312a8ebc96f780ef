// Probe: are the predicate outputs of __vibmax_s16x2 what the CUDA header documents ((a >= b) per half) for every
// operand order ptxas may pick -- including a constant-zero first operand?  Exhaustive over one half (65536 x 9 values).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(unsigned long long *bad, uint32_t zr /* 0 at run time, unknown to ptxas */)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;           // 0..65535: low half; high half = ~x
	const uint32_t a = (x & 0xffffu) | ((~x & 0xffffu) << 16);
	const int16_t others[9] = {0, 1, -1, 256, -256, 32767, -32768, (int16_t)x, (int16_t)(x + 1)};
	for (int k = 0; k < 9; ++k) {
		const uint32_t b = ((uint16_t)others[k]) | ((uint32_t)(uint16_t)others[(k + 3) % 9] << 16);
		bool h, l;
		// form 1: variable, variable
		uint32_t m = __vibmax_s16x2(a, b, &h, &l);
		int16_t al = (int16_t)(a & 0xffff), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xffff), bh = (int16_t)(b >> 16);
		uint32_t em = (uint16_t)(al > bl ? al : bl) | ((uint32_t)(uint16_t)(ah > bh ? ah : bh) << 16);
		if (m != em || l != (al >= bl) || h != (ah >= bh)) atomicAdd(&bad[0], 1ull);
		// form 2: zero first
		m = __vibmax_s16x2(0u, a, &h, &l);
		em = (uint16_t)(al > 0 ? al : 0) | ((uint32_t)(uint16_t)(ah > 0 ? ah : 0) << 16);
		if (m != em || l != (0 >= al) || h != (0 >= ah)) atomicAdd(&bad[1], 1ull);
		// form 3: zero second
		m = __vibmax_s16x2(a, 0u, &h, &l);
		if (m != em || l != (al >= 0) || h != (ah >= 0)) atomicAdd(&bad[2], 1ull);
		// form 5: run-time zero first (what the packed kernel uses for  bit = !(0 >= a))
		m = __vibmax_s16x2(zr, a, &h, &l);
		if (m != em || l != (0 >= al) || h != (0 >= ah)) atomicAdd(&bad[4], 1ull);
		// form 6: predicates only, both variable, result unused
		(void)__vibmax_s16x2(a ^ zr, m, &h, &l);
		{ int16_t ml = (int16_t)(m & 0xffff), mh = (int16_t)(m >> 16); if (l != (al >= ml) || h != (ah >= mh)) atomicAdd(&bad[5], 1ull); }
		// form 4: result unused (predicates only)
		(void)__vibmax_s16x2(b, a, &h, &l);
		if (l != (bl >= al) || h != (bh >= ah)) atomicAdd(&bad[3], 1ull);
	}
}
int main()
{
	unsigned long long *d, h[6];
	cudaMalloc(&d, sizeof(h)); cudaMemset(d, 0, sizeof(h));
	probe<<<256, 256>>>(d, 0u);
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	printf("vibmax_s16x2 mismatches: var,var %llu | (0,a) %llu | (a,0) %llu | preds-only(const) %llu | (zr,a) %llu | preds-only(var) %llu   (%s)\n", h[0], h[1], h[2], h[3], h[4], h[5],
	       cudaGetErrorString(cudaGetLastError()));
	return 0;
}
