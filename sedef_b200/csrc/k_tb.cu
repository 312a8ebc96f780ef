// traceback + statistics kernels, input preparation and result gather
#include "kernels.h"
#include "extz_tb.cuh"
#include "extz_io.cuh"

namespace extz {

cudaError_t k_traceback_launch(const TbLaunch &L, bool warp_per_pair, bool stats, cudaStream_t st)
{
	if (L.n <= 0) return cudaSuccess;
	if (warp_per_pair) {
		const int grid = (L.n + 3) / 4;
		if (stats) extz_traceback_warp_kernel<true><<<grid, 128, 0, st>>>(L);
		else extz_traceback_warp_kernel<false><<<grid, 128, 0, st>>>(L);
	} else {
		const int grid = (L.n + 127) / 128;
		if (stats) extz_traceback_kernel<true><<<grid, 128, 0, st>>>(L);
		else extz_traceback_kernel<false><<<grid, 128, 0, st>>>(L);
	}
	return cudaGetLastError();
}
cudaError_t k_stats_from_cigar_launch(const CigarStatsLaunch &L, cudaStream_t st)
{
	if (L.n <= 0) return cudaSuccess;
	sd_stats_from_cigar_kernel<<<(L.n + 3) / 4, 128, 0, st>>>(L);                // one warp per alignment
	return cudaGetLastError();
}
static inline int io_grid(size_t nvec)
{
	const size_t want = (nvec + 255) / 256;
	return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));      // grid-stride loops: a multiple of the SM count is enough
}
cudaError_t k_encode_launch(const uint8_t *raw, uint8_t *codes, size_t nbytes, cudaStream_t st)
{
	if (nbytes == 0) return cudaSuccess;
	encode_kernel<<<io_grid(nbytes >> 4), 256, 0, st>>>(raw, codes, nbytes);
	return cudaGetLastError();
}
cudaError_t k_check_symbols_launch(const PairDesc *pairs, int n, const uint8_t *codes, int limit, int *flag, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	check_symbols_kernel<<<io_grid(((size_t)n + 7) / 8 * 256), 256, 0, st>>>(pairs, n, codes, (uint32_t)limit, flag);
	return cudaGetLastError();
}
cudaError_t k_fill_reset_launch(uint64_t *ez_out, int n, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	fill_reset_kernel<<<(n + 255) / 256, 256, 0, st>>>(ez_out, n);
	return cudaGetLastError();
}
cudaError_t k_gather_launch(const GatherLaunch &L, cudaStream_t st)
{
	if (L.n <= 0) return cudaSuccess;
	gather_kernel<<<(L.n + 255) / 256, 256, 0, st>>>(L);
	return cudaGetLastError();
}

} // namespace extz
