// Probe: does the fused VIADDMNMX.S16x2 (__viaddmax_s16x2) wrap its 16-bit add exactly like add.s16x2 followed by
// max.s16x2 (the PTX it is defined as)?  The packed DP kernel relies on int8-in-the-top-byte wrap-around.
// Also checks the unsigned __dp4a byte pick of the lazy-H update and the packed min/max/add primitives against scalar code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(unsigned long long *bad, uint32_t zr)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;           // 0..65535
	for (uint32_t y = 0; y < 65536; y += 251) {
		const uint32_t a = x | ((x ^ 0x5a5au) << 16), b = (y ^ zr) | (((y * 7u) & 0xffffu) << 16);
		const int16_t al = (int16_t)(a & 0xffff), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xffff), bh = (int16_t)(b >> 16);
		const int16_t sl = (int16_t)(uint16_t)((uint16_t)al + (uint16_t)bl), sh = (int16_t)(uint16_t)((uint16_t)ah + (uint16_t)bh);
		uint32_t got = __viaddmax_s16x2(a, b, 0u);
		uint32_t exp = (uint16_t)(sl > 0 ? sl : 0) | ((uint32_t)(uint16_t)(sh > 0 ? sh : 0) << 16);
		if (got != exp) atomicAdd(&bad[0], 1ull);
		got = __vadd2(a, b); exp = (uint16_t)sl | ((uint32_t)(uint16_t)sh << 16);
		if (got != exp) atomicAdd(&bad[1], 1ull);
		got = __vmaxs2(a, b); exp = (uint16_t)(al > bl ? al : bl) | ((uint32_t)(uint16_t)(ah > bh ? ah : bh) << 16);
		if (got != exp) atomicAdd(&bad[2], 1ull);
		const uint16_t ual = a & 0xffff, uah = a >> 16, ubl = b & 0xffff, ubh = b >> 16;
		got = __vmaxu2(a, b); exp = (uint32_t)(ual > ubl ? ual : ubl) | ((uint32_t)(uah > ubh ? uah : ubh) << 16);
		if (got != exp) atomicAdd(&bad[3], 1ull);
		got = __vminu2(a, b); exp = (uint32_t)(ual < ubl ? ual : ubl) | ((uint32_t)(uah < ubh ? uah : ubh) << 16);
		if (got != exp) atomicAdd(&bad[4], 1ull);
		// unsigned dp4a: ZERO-extended byte 1 / byte 3 added to an accumulator (H[t] += v8[t], v8 = uint8_t*); the accumulator
		// is a two's-complement int32 carried as uint32
		const uint32_t acc = y * 2654435761u;
		if (__dp4a(a, 0x00000100u, acc) != acc + ((a >> 8) & 0xffu)) atomicAdd(&bad[5], 1ull);
		if (__dp4a(a, 0x01000000u, acc) != acc + (a >> 24)) atomicAdd(&bad[5], 1ull);
	}
}
int main()
{
	unsigned long long *d, h[6];
	cudaMalloc(&d, sizeof(h)); cudaMemset(d, 0, sizeof(h));
	probe<<<256, 256>>>(d, 0u);
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	printf("mismatches: viaddmax %llu | vadd2 %llu | vmaxs2 %llu | vmaxu2 %llu | vminu2 %llu | dp4a %llu   (%s)\n",
	       h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(cudaGetLastError()));
	return 0;
}
