// Positive k-mer hash set and the negative sampler (main.py:361-459 generate_negative, :345-346
// neighbor_check, utils.py:75-97 build_hash).  The reference tests membership in a Bloom filter
// (pybloom_live, 1e-3 false positives); here membership is EXACT: an open-addressing table of 16-byte
// slots holding the packed k-mer, probed linearly.  Restated bit-for-bit (same splitmix64 streams,
// exact Python set) in oracle/sampler_oracle.py.
#include "common.cuh"
#include "rowwise.cuh"

namespace matcha {
namespace {

constexpr int kMaxW = 6;          // widest hyperedge the 128-bit key holds (21 bits per id)
constexpr int64_t kMaxId = (1ll << 21) - 1;

struct Key { unsigned long long lo, hi; };

// ids sorted ascending, zero padded; e[0] >= 1 so lo != 0 for every real k-mer
__device__ __forceinline__ Key pack_key(const int64_t* e, int L) {
  unsigned long long v[kMaxW] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < kMaxW; ++i) if (i < L) v[i] = (unsigned long long)e[i];
  Key k;
  k.lo = v[0] | (v[1] << 21) | (v[2] << 42);
  k.hi = v[3] | (v[4] << 21) | (v[5] << 42);
  return k;
}
__device__ __forceinline__ uint64_t hash_key(Key k) { return splitmix64(k.lo ^ splitmix64(k.hi)); }

__device__ __forceinline__ bool table_contains(const ulonglong2* __restrict__ tab, uint64_t mask, Key k) {
  uint64_t slot = hash_key(k) & mask;
  for (uint64_t probe = 0; probe <= mask; ++probe) {
    const ulonglong2 s = __ldg(tab + slot);
    if (s.x == 0ull) return false;
    if (s.x == k.lo && s.y == k.hi) return true;
    slot = (slot + 1) & mask;
  }
  return false;
}

__global__ void hashset_insert_kernel(unsigned long long* tab, uint64_t mask, const int64_t* __restrict__ kmers, int64_t n,
                                      int L, int* overflow) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const Key k = pack_key(kmers + i * L, L);
    if (k.lo == 0ull) continue;   // empty row
    uint64_t slot = hash_key(k) & mask;
    bool done = false;
    for (uint64_t probe = 0; probe <= mask && !done; ++probe) {
      const unsigned long long old = atomicCAS(tab + 2 * slot, 0ull, k.lo);
      if (old == 0ull) { tab[2 * slot + 1] = k.hi; done = true; }
      else slot = (slot + 1) & mask;
    }
    if (!done) atomicExch(overflow, 1);
  }
}
__global__ void hashset_contains_kernel(const ulonglong2* __restrict__ tab, uint64_t mask, const int64_t* __restrict__ kmers,
                                        int64_t n, int L, uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const Key k = pack_key(kmers + i * L, L);
    out[i] = (k.lo != 0ull && table_contains(tab, mask, k)) ? 1 : 0;
  }
}

// One thread per negative.  Semantics of generate_negative:
//   positions to corrupt: every position independently with probability 1/2, conditioned on at least
//   one (== Binomial(k, 1/2) | >0 count + uniform subset, main.py:371-372,389), FIXED across retries;
//   each candidate round replaces them by a uniform bin of the SAME chromosome (main.py:402-407),
//   sorts, rejects duplicates (main.py:410-414), adjacent gaps <= min_dis (:416-421) and members of the
//   positive set (:392).  The reference retries forever; we stop after max_rounds and flag the row.
__global__ void neg_sample_kernel(const ulonglong2* __restrict__ tab, uint64_t mask, const int64_t* __restrict__ pos,
                                  int64_t P, int L, int neg_num, const ChromMeta cm, int min_dis, uint64_t seed,
                                  uint64_t step, int max_rounds, int64_t* __restrict__ neg, uint8_t* __restrict__ valid,
                                  int32_t* __restrict__ rounds_used) {
  const int64_t total = P * neg_num;
  const uint64_t base = splitmix64(seed ^ splitmix64(step));
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t* prow = pos + (g / neg_num) * L;
    int64_t p[kMaxW], t[kMaxW];
    int k = 0;
#pragma unroll
    for (int i = 0; i < kMaxW; ++i) { p[i] = (i < L) ? prow[i] : 0; if (p[i] != 0) k = i + 1; }
    const uint64_t key = splitmix64(base + (uint64_t)g);
    uint64_t ctr = 0;
    uint32_t cmask = 0;
    for (int tries = 0; tries < 64 && cmask == 0; ++tries) cmask = (uint32_t)(splitmix64(key + ctr++) & ((1ull << k) - 1ull));
    if (cmask == 0) cmask = 1;
    // chromosome range of every position to be replaced (depends on the positive only)
    int64_t cs[kMaxW], ce[kMaxW];
#pragma unroll
    for (int i = 0; i < kMaxW; ++i) {
      cs[i] = 0; ce[i] = 0;
      if (i < k && ((cmask >> i) & 1u)) {
        for (int c = 0; c < cm.n; ++c)
          if (p[i] >= cm.start[c] && p[i] < cm.end[c]) { cs[i] = cm.start[c]; ce[i] = cm.end[c]; }
      }
    }
    bool accepted = false;
    int round = 0;
    for (; round < max_rounds && !accepted; ++round) {
#pragma unroll
      for (int i = 0; i < kMaxW; ++i) {
        t[i] = p[i];
        if (i < k && ((cmask >> i) & 1u) && ce[i] > cs[i]) {
          const uint64_t u = splitmix64(key + ctr++) >> 32;
          t[i] = cs[i] + (int64_t)((u * (uint64_t)(ce[i] - cs[i])) >> 32);
        }
      }
      // insertion sort of the k live entries
#pragma unroll
      for (int i = 1; i < kMaxW; ++i) {
        if (i < k) {
          const int64_t v = t[i];
          int j = i - 1;
          while (j >= 0 && t[j] > v) { t[j + 1] = t[j]; --j; }
          t[j + 1] = v;
        }
      }
      bool ok = true;
#pragma unroll
      for (int i = 0; i + 1 < kMaxW; ++i)
        if (i + 1 < k) { const int64_t gap = t[i + 1] - t[i]; if (gap == 0 || gap <= min_dis) ok = false; }
      if (ok && !table_contains(tab, mask, pack_key(t, k))) accepted = true;
    }
    int64_t* out = neg + g * L;
    for (int i = 0; i < L; ++i) out[i] = accepted ? (i < k ? t[i] : 0) : p[i];
    if (valid) valid[g] = accepted ? 1 : 0;
    if (rounds_used) rounds_used[g] = round;
  }
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

int matcha_hashset_insert(void* table, int64_t capacity, const int64_t* kmers, int64_t n, int32_t width, void* stream) {
  MATCHA_REQUIRE(table && kmers && n >= 0, "hashset_insert: NULL argument");
  MATCHA_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0, "hashset capacity must be a power of two");
  MATCHA_REQUIRE(width >= 1 && width <= kMaxW, "hashset width %d unsupported (1..%d)", width, kMaxW);
  MATCHA_REQUIRE(n * 2 <= capacity, "hashset load factor above 0.5 (n=%lld capacity=%lld)", (long long)n, (long long)capacity);
  if (n == 0) return MATCHA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  // slot `capacity` (one past the table proper) is the status word: set to 1 if an insert found no free slot
  hashset_insert_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<unsigned long long*>(table), (uint64_t)capacity - 1, kmers, n, width,
      reinterpret_cast<int*>(reinterpret_cast<unsigned long long*>(table) + 2 * capacity));
  MATCHA_CHECK_LAUNCH("hashset_insert");
  return MATCHA_OK;
}

int matcha_hashset_contains(const void* table, int64_t capacity, const int64_t* kmers, int64_t n, int32_t width,
                            uint8_t* out, void* stream) {
  MATCHA_REQUIRE(table && kmers && out, "hashset_contains: NULL argument");
  MATCHA_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0, "hashset capacity must be a power of two");
  MATCHA_REQUIRE(width >= 1 && width <= kMaxW, "hashset width %d unsupported (1..%d)", width, kMaxW);
  if (n == 0) return MATCHA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  hashset_contains_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const ulonglong2*>(table),
                                                                        (uint64_t)capacity - 1, kmers, n, width, out);
  MATCHA_CHECK_LAUNCH("hashset_contains");
  return MATCHA_OK;
}

int matcha_neg_sample(const void* table, int64_t capacity, const int64_t* pos, int64_t P, int32_t L, int32_t neg_num,
                      const int64_t* chrom_start, const int64_t* chrom_end, int32_t n_chrom, int32_t min_dis,
                      uint64_t seed, uint64_t step, int32_t max_rounds, int64_t* neg, uint8_t* valid,
                      int32_t* rounds_used, void* stream) {
  MATCHA_REQUIRE(table && pos && neg && chrom_start && chrom_end, "neg_sample: NULL argument");
  MATCHA_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0, "hashset capacity must be a power of two");
  MATCHA_REQUIRE(L >= 1 && L <= kMaxW, "neg_sample: width %d unsupported (1..%d)", L, kMaxW);
  MATCHA_REQUIRE(n_chrom >= 1 && n_chrom <= MATCHA_MAX_CHROM, "neg_sample: n_chrom out of range");
  MATCHA_REQUIRE(neg_num >= 1 && max_rounds >= 1, "neg_sample: neg_num / max_rounds must be positive");
  ChromMeta cm;
  cm.n = n_chrom;
  for (int c = 0; c < n_chrom; ++c) {
    cm.start[c] = chrom_start[c]; cm.end[c] = chrom_end[c];   // HOST arrays
    MATCHA_REQUIRE(chrom_end[c] - 1 <= kMaxId, "node ids above 2^21-1 do not fit the 128-bit k-mer key");
  }
  const int64_t total = P * neg_num;
  if (total == 0) return MATCHA_OK;
  int64_t blocks = (total + 127) / 128;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  prof_begin(P_SAMPLER, (cudaStream_t)stream);
  neg_sample_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const ulonglong2*>(table),
                                                                  (uint64_t)capacity - 1, pos, P, L, neg_num, cm, min_dis,
                                                                  seed, step, max_rounds, neg, valid, rounds_used);
  MATCHA_CHECK_LAUNCH("neg_sample");
  prof_end(P_SAMPLER, 1, (cudaStream_t)stream);
  return MATCHA_OK;
}

}  // extern "C"
