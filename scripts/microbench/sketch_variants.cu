// Micro-benchmark (development tool, not part of the library): where should the integer work of mm_sketch_two's inner loop
// (roll the forward / reverse k-mer, hash64 of the canonical one, keep the minimum) run?  The SASS of k_classify_sketch_packed
// has ~37 ALU-pipe instructions (LOP3, SHF, ISETP, SEL) and ~10 FMA-pipe instructions (IMAD) per position; either pipe takes one
// warp instruction every two cycles per SM sub-partition (B300_MICROARCH.md), so the loop is bound by the ALU pipe (74 cycles
// per position per warp, issue slots ~65 % busy: what ncu reports).  The variants move shifts onto the FMA pipe as multiplications
// by powers of two held in registers (ptxas cannot turn those back into shifts):
//   HV 0  hash64 as in mcb_hash64_wide (shifts: SHF + LOP3)
//   HV 1  every x ^= x >> s stage as IMAD + IMAD.HI (+ one LOP3 per half)
//   HV 2  only the low half's funnel shift as IMAD + IMAD.HI
//   RV 0  roll with funnel shifts;  RV 1  roll with IMAD / IMAD.HI;  RV 2  only the reverse k-mer's high half that way
// Every variant must produce the same (hash, position) per read: a checksum is printed.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o sketch_variants sketch_variants.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

struct Mul { uint32_t mh, m24, m14, m28, u24, u14, u28, sh2, four, p30, negs1, t3; int s1; };

__device__ __forceinline__ uint32_t madhi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

template <int HV>
__device__ __forceinline__ uint64_t hash_v(uint32_t lo, uint32_t hi, const Mul &M)
{
	const uint32_t mh = M.mh;
	uint64_t t = (uint64_t)lo * 0x1FFFFFu + 0xFFFFFFFFFFFFFFFFull;
	uint32_t raw = hi * 0x1FFFFFu + (uint32_t)(t >> 32); lo = (uint32_t)t;
#define STAGE(S, MS, US) \
	if (HV == 0) { hi = raw & mh; lo ^= __funnelshift_r(lo, hi, S); hi ^= hi >> S; } \
	else if (HV == 1) { const uint32_t tl = madhi(lo, MS, raw * MS); const uint32_t u = __umulhi(raw * M.sh2, US); lo ^= tl; hi = (raw & mh) ^ u; } \
	else { hi = raw & mh; lo ^= madhi(lo, MS, raw * MS); hi ^= hi >> S; }
	STAGE(24, M.m24, M.u24)
	t = (uint64_t)lo * 265u; raw = hi * 265u + (uint32_t)(t >> 32); lo = (uint32_t)t;
	STAGE(14, M.m14, M.u14)
	t = (uint64_t)lo * 21u; raw = hi * 21u + (uint32_t)(t >> 32); lo = (uint32_t)t;
	STAGE(28, M.m28, M.u28)
	t = (uint64_t)lo * 0x80000001u; hi = (hi * 0x80000001u + (uint32_t)(t >> 32)) & mh; lo = (uint32_t)t;
#undef STAGE
	return (uint64_t)hi << 32 | lo;
}

template <int HV, int RV>
__global__ void __launch_bounds__(128) k_sketch(const uint64_t *__restrict__ rows, uint64_t n, int L, int k, Mul M, uint64_t *__restrict__ out_x, int *__restrict__ out_p)
{
	__shared__ uint64_t sp[128][5];
	const int t = threadIdx.x;
	const uint64_t i = (uint64_t)blockIdx.x * 128 + t;
	if (i >= n) return;
	const uint4 *s4 = (const uint4*)(rows + i * 4);
	for (int v = 0; v < 2; ++v) { const uint4 q = s4[v]; sp[t][2 * v] = (uint64_t)q.y << 32 | q.x; sp[t][2 * v + 1] = (uint64_t)q.w << 32 | q.z; }
	const uint32_t mh = M.mh;
	const int s1 = M.s1;
	const uint32_t t3 = M.t3;
	uint32_t flo = 0, fhi = 0, rlo = 0, rhi = 0, blo = ~0u, bhi = ~0u;
	int bp = -1;
#define ROLL() { const uint32_t c = x & 3u; x >>= 2; \
	if (RV == 1) { fhi = madhi(flo, M.four, fhi * M.four) & mh; flo = flo * 4u + c; rlo = madhi(rlo, M.p30, rhi * M.p30); rhi = madhi(rhi, M.p30, c * M.negs1 + t3); } \
	else if (RV == 2) { fhi = __funnelshift_l(flo, fhi, 2) & mh; flo = flo * 4u + c; rlo = __funnelshift_r(rlo, rhi, 2); rhi = madhi(rhi, M.p30, c * M.negs1 + t3); } \
	else { fhi = __funnelshift_l(flo, fhi, 2) & mh; flo = flo * 4u + c; rlo = __funnelshift_r(rlo, rhi, 2); rhi = (rhi >> 2) | ((c << s1) ^ t3); } }
	for (int h = 0; h * 16 < L; ++h) {
		uint32_t x = (uint32_t)(sp[t][h >> 1] >> (32 * (h & 1)));
		const int base = h * 16;
		const int lim = min(16, L - base);
		const int split = max(0, min(lim, k - 1 - base));
		for (int j = 0; j < split; ++j) ROLL();
		for (int j = split; j < lim; ++j) {
			ROLL();
			const bool fwd = ((uint64_t)fhi << 32 | flo) < ((uint64_t)rhi << 32 | rlo);
			const uint64_t hv = hash_v<HV>(fwd ? flo : rlo, fwd ? fhi : rhi, M);
			const bool better = hv < ((uint64_t)bhi << 32 | blo);
			blo = better ? (uint32_t)hv : blo; bhi = better ? (uint32_t)(hv >> 32) : bhi; bp = better ? base + j : bp;
		}
	}
#undef ROLL
	out_x[i] = (uint64_t)bhi << 32 | blo; out_p[i] = bp;
}


// ---- "top-aligned" arithmetic: the 2k-bit k-mers and hashes live in the TOP 2k bits of the 64-bit pair (value << (64 - 2k)).
// Multiplication modulo 2^(2k) is then plain wrap-around modulo 2^64, so the four "& mask" after hash64's multiplications
// disappear; x ^= x >> s needs the shifted-in low bits cleared, which folds into the XOR (one LOP3: a ^ (b & c)).  Order is
// preserved by the shift, so minima and the f < r test are unchanged.
//   TV 0  shifts on the ALU pipe;  TV 1  the high half's x >> s as IMAD.HI;  TV 2  also the reverse k-mer's high half as IMAD.HI;
//   TV 3  also the first stage's low-half funnel shift as IMAD + IMAD.HI
struct MulT { uint32_t lm, csh, m24, m14, m28, p30, neg30, sub1lo, sub1hi; int sh; };
template <int TV>
__device__ __forceinline__ uint64_t hash_ta(uint32_t lo, uint32_t hi, const MulT &M)
{
	uint64_t t = (uint64_t)lo * 0x1FFFFFu + ((uint64_t)M.sub1hi << 32 | M.sub1lo);
	hi = hi * 0x1FFFFFu + (uint32_t)(t >> 32); lo = (uint32_t)t;
#define TSTAGE(S, MS, LOFMA) { \
	const uint32_t fl = (TV >= 3 && LOFMA) ? madhi(lo, MS, hi * MS) : __funnelshift_r(lo, hi, S); \
	const uint32_t u = TV >= 1 ? __umulhi(hi, MS) : hi >> S; \
	lo ^= fl & M.lm; hi ^= u; }
	TSTAGE(24, M.m24, true)
	t = (uint64_t)lo * 265u; hi = hi * 265u + (uint32_t)(t >> 32); lo = (uint32_t)t;
	TSTAGE(14, M.m14, false)
	t = (uint64_t)lo * 21u; hi = hi * 21u + (uint32_t)(t >> 32); lo = (uint32_t)t;
	TSTAGE(28, M.m28, false)
	t = (uint64_t)lo * 0x80000001u; hi = hi * 0x80000001u + (uint32_t)(t >> 32); lo = (uint32_t)t;
#undef TSTAGE
	return (uint64_t)hi << 32 | lo;
}

template <int TV>
__global__ void __launch_bounds__(128) k_sketch_ta(const uint64_t *__restrict__ rows, uint64_t n, int L, int k, MulT M, uint64_t *__restrict__ out_x, int *__restrict__ out_p)
{
	__shared__ uint64_t sp[128][5];
	const int t = threadIdx.x;
	const uint64_t i = (uint64_t)blockIdx.x * 128 + t;
	if (i >= n) return;
	const uint4 *s4 = (const uint4*)(rows + i * 4);
	for (int v = 0; v < 2; ++v) { const uint4 q = s4[v]; sp[t][2 * v] = (uint64_t)q.y << 32 | q.x; sp[t][2 * v + 1] = (uint64_t)q.w << 32 | q.z; }
	uint32_t flo = 0, fhi = 0, rlo = 0, rhi = 0, blo = ~0u, bhi = ~0u;
	int bp = -1;
#define TROLL() { const uint32_t c = x & 3u; x >>= 2; \
	fhi = __funnelshift_l(flo, fhi, 2); flo = flo * 4u + c * M.csh; \
	rlo = __funnelshift_r(rlo, rhi, 2) & M.lm; \
	const uint32_t X = c * M.neg30 + 0xC0000000u; \
	rhi = TV >= 2 ? madhi(rhi, M.p30, X) : ((rhi >> 2) | X); }
	for (int h = 0; h * 16 < L; ++h) {
		uint32_t x = (uint32_t)(sp[t][h >> 1] >> (32 * (h & 1)));
		const int base = h * 16;
		const int lim = min(16, L - base);
		const int split = max(0, min(lim, k - 1 - base));
		for (int j = 0; j < split; ++j) TROLL();
		for (int j = split; j < lim; ++j) {
			TROLL();
			const bool fwd = ((uint64_t)fhi << 32 | flo) < ((uint64_t)rhi << 32 | rlo);
			const uint64_t hv = hash_ta<TV>(fwd ? flo : rlo, fwd ? fhi : rhi, M);
			const bool better = hv < ((uint64_t)bhi << 32 | blo);
			blo = better ? (uint32_t)hv : blo; bhi = better ? (uint32_t)(hv >> 32) : bhi; bp = better ? base + j : bp;
		}
	}
#undef TROLL
	out_x[i] = bp < 0 ? ~0ull : ((uint64_t)bhi << 32 | blo) >> M.sh; out_p[i] = bp;
}

template <int TV>
static void run_ta(const char *name, const uint64_t *d_rows, uint64_t n, int L, int k, MulT M, uint64_t *d_x, int *d_p, std::vector<uint64_t> &hx, std::vector<int> &hp)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const unsigned grid = (unsigned)((n + 127) / 128);
	for (int w = 0; w < 3; ++w) k_sketch_ta<TV><<<grid, 128>>>(d_rows, n, L, k, M, d_x, d_p);
	cudaEventRecord(a);
	const int reps = 10;
	for (int r = 0; r < reps; ++r) k_sketch_ta<TV><<<grid, 128>>>(d_rows, n, L, k, M, d_x, d_p);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms = 0; cudaEventElapsedTime(&ms, a, b);
	cudaMemcpy(hx.data(), d_x, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hp.data(), d_p, n * 4, cudaMemcpyDeviceToHost);
	uint64_t cs = 0;
	for (uint64_t i = 0; i < n; ++i) cs = cs * 1000003ull + hx[i] * 31ull + (uint64_t)hp[i];
	cudaError_t e = cudaGetLastError();
	printf("%-28s %8.4f ms per %llu reads  (%.1f M reads/s)  checksum %016llx  %s\n", name, ms / reps, (unsigned long long)n, n / (ms / reps) / 1e3, (unsigned long long)cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int HV, int RV>
static void run(const char *name, const uint64_t *d_rows, uint64_t n, int L, int k, Mul M, uint64_t *d_x, int *d_p, std::vector<uint64_t> &hx, std::vector<int> &hp)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const unsigned grid = (unsigned)((n + 127) / 128);
	for (int w = 0; w < 3; ++w) k_sketch<HV, RV><<<grid, 128>>>(d_rows, n, L, k, M, d_x, d_p);
	cudaEventRecord(a);
	const int reps = 10;
	for (int r = 0; r < reps; ++r) k_sketch<HV, RV><<<grid, 128>>>(d_rows, n, L, k, M, d_x, d_p);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms = 0; cudaEventElapsedTime(&ms, a, b);
	cudaMemcpy(hx.data(), d_x, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hp.data(), d_p, n * 4, cudaMemcpyDeviceToHost);
	uint64_t cs = 0;
	for (uint64_t i = 0; i < n; ++i) cs = cs * 1000003ull + hx[i] * 31ull + (uint64_t)hp[i];
	cudaError_t e = cudaGetLastError();
	printf("%-28s %8.4f ms per %llu reads  (%.1f M reads/s)  checksum %016llx  %s\n", name, ms / reps, (unsigned long long)n, n / (ms / reps) / 1e3, (unsigned long long)cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char **argv)
{
	const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 10000000ull;
	const int L = 100, k = argc > 2 ? atoi(argv[2]) : 31;
	std::vector<uint64_t> rows(n * 4);
	uint64_t s = 0x9E3779B97F4A7C15ull;
	for (auto &w : rows) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w = s; }
	for (uint64_t i = 0; i < n; ++i) rows[i * 4 + 3] &= 0xFFull;         // bases 96..99 only
	uint64_t *d_rows, *d_x; int *d_p;
	cudaMalloc(&d_rows, n * 32); cudaMalloc(&d_x, n * 8); cudaMalloc(&d_p, n * 4);
	cudaMemcpy(d_rows, rows.data(), n * 32, cudaMemcpyHostToDevice);
	Mul M;
	const int hb = 2 * k - 32;                                            // bits of the high half
	M.mh = (1u << hb) - 1u; M.m24 = 1u << 8; M.m14 = 1u << 18; M.m28 = 1u << 4;
	auto up = [&](int sh) { const int e = hb - sh; return e >= 0 ? (1u << e) : 0u; };
	M.u24 = up(24); M.u14 = up(14); M.u28 = up(28); M.sh2 = 1u << (32 - hb);
	M.four = 4u; M.p30 = 1u << 30; M.s1 = 2 * (k - 1) - 32; M.negs1 = 0u - (1u << M.s1); M.t3 = 3u << M.s1;
	if (hb < 28) { printf("the IMAD forms need 2k - 32 >= 28 (k >= 30)\n"); return 1; }
	std::vector<uint64_t> hx(n); std::vector<int> hp(n);
	run<0, 0>("HV0 RV0 (library today)", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<1, 0>("HV1 RV0", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<2, 0>("HV2 RV0", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<0, 1>("HV0 RV1", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<0, 2>("HV0 RV2", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<1, 1>("HV1 RV1", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<1, 2>("HV1 RV2", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<2, 1>("HV2 RV1", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	run<2, 2>("HV2 RV2", d_rows, n, L, k, M, d_x, d_p, hx, hp);
	MulT T;
	T.sh = 64 - 2 * k; T.lm = ~((1u << T.sh) - 1u); T.csh = 1u << T.sh; T.m24 = 1u << 8; T.m14 = 1u << 18; T.m28 = 1u << 4; T.p30 = 1u << 30; T.neg30 = 0u - (1u << 30);
	{ const uint64_t sub = 0ull - (1ull << T.sh); T.sub1lo = (uint32_t)sub; T.sub1hi = (uint32_t)(sub >> 32); }
	run_ta<0>("top-aligned TV0", d_rows, n, L, k, T, d_x, d_p, hx, hp);
	run_ta<1>("top-aligned TV1", d_rows, n, L, k, T, d_x, d_p, hx, hp);
	run_ta<2>("top-aligned TV2", d_rows, n, L, k, T, d_x, d_p, hx, hp);
	run_ta<3>("top-aligned TV3", d_rows, n, L, k, T, d_x, d_p, hx, hp);
	return 0;
}
