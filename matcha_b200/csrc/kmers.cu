// k-mer enumeration and counting: the producer of the hot path's inputs,
// generate_kmers.py:8-69 (build_dict) and :86-141, as two kernels.
//   count:   one thread per (cluster, k-subset).  Clusters are CSR rows of unique ascending node ids; the host supplies
//            the prefix sums of C(n_c, k) over the eligible clusters (k <= n_c <= max_cluster_size, :88-91).  A thread
//            finds its cluster by binary search, unranks its subset (lexicographic combinatorial number system), keeps
//            it iff every adjacent gap exceeds min_distance (:16-17 anchor rule + :23-32) and bumps the k-mer's counter:
//            find-or-insert into an open-addressing table of 16-byte packed keys with ONE 128-bit compare-and-swap
//            (no half-written slot can ever be observed), then an atomic add on the slot's count.
//   collect: slots whose count reaches min_freq_cutoff (:40) are appended to the output (order is irrelevant: the
//            reference's own order is process-pool completion order; the host sorts).
// Integer work, bit-exact against oracle/kmer_oracle.py and the reference's own output (tests/golden/kmer_small.npz).
#include "common.cuh"

namespace matcha {
namespace {

constexpr int kKmerMaxK = 6;                 // 6 ids x 21 bits in 128 bits (same packing as the positive hash set)
constexpr int kKmerMaxN = 64;                // largest cluster the binomial table covers
constexpr int64_t kKmerMaxId = (1ll << 21) - 1;

struct BinomTable { unsigned long long c[kKmerMaxN + 1][kKmerMaxK + 1]; };

struct Key128 { unsigned long long lo, hi; };

__device__ __forceinline__ uint64_t hash128(Key128 k) { return splitmix64(k.lo ^ splitmix64(k.hi)); }

// 128-bit compare-and-swap against the empty slot (0, 0); returns the previous contents
__device__ __forceinline__ Key128 cas128_empty(unsigned long long* slot, Key128 val) {
  Key128 old;
  asm volatile(
      "{\n\t"
      ".reg .b128 cmp, val, old;\n\t"
      "mov.b128 cmp, {%2, %3};\n\t"
      "mov.b128 val, {%4, %5};\n\t"
      "atom.global.cas.b128 old, [%6], cmp, val;\n\t"
      "mov.b128 {%0, %1}, old;\n\t"
      "}\n"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(0ull), "l"(0ull), "l"(val.lo), "l"(val.hi), "l"(slot)
      : "memory");
  return old;
}

__global__ void kmer_count_kernel(const int64_t* __restrict__ members, const int64_t* __restrict__ offsets, int64_t n_clusters,
                                  const int64_t* __restrict__ work_prefix, int k, int min_dis, const BinomTable bt,
                                  unsigned long long* __restrict__ table, uint64_t mask, int32_t* __restrict__ counts,
                                  int* __restrict__ status) {
  const int64_t total = work_prefix[n_clusters];
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    // cluster c with work_prefix[c] <= w < work_prefix[c + 1]
    int64_t lo = 0, hi = n_clusters;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(work_prefix + mid) <= w) lo = mid; else hi = mid;
    }
    const int64_t c = lo;
    unsigned long long r = (unsigned long long)(w - __ldg(work_prefix + c));
    const int64_t b = __ldg(offsets + c);
    const int n = (int)(__ldg(offsets + c + 1) - b);
    // unrank: position j takes the smallest element e (after the previous one) whose block of C(n - e - 1, k - j - 1)
    // subsets contains the remaining rank
    int64_t ids[kKmerMaxK];
    int e = 0;
    bool keep = n <= kKmerMaxN;
#pragma unroll
    for (int j = 0; j < kKmerMaxK; ++j) {
      ids[j] = 0;
      if (j < k && keep) {
        while (e < n) {
          const unsigned long long blockc = bt.c[n - e - 1][k - j - 1];
          if (r < blockc) break;
          r -= blockc;
          ++e;
        }
        if (e >= n) { keep = false; } else { ids[j] = __ldg(members + b + e); ++e; }
      }
    }
    if (!keep) { atomicExch(status, 2); continue; }
#pragma unroll
    for (int j = 0; j + 1 < kKmerMaxK; ++j)
      if (j + 1 < k && ids[j + 1] - ids[j] <= (int64_t)min_dis) keep = false;
    if (!keep) continue;
    if (ids[0] < 1 || ids[k - 1] > kKmerMaxId) { atomicExch(status, 3); continue; }
    Key128 key;
    key.lo = (unsigned long long)ids[0] | ((unsigned long long)ids[1] << 21) | ((unsigned long long)ids[2] << 42);
    key.hi = (unsigned long long)ids[3] | ((unsigned long long)ids[4] << 21) | ((unsigned long long)ids[5] << 42);
    uint64_t slot = hash128(key) & mask;
    bool done = false;
    for (uint64_t probe = 0; probe <= mask && !done; ++probe) {
      const Key128 old = cas128_empty(table + 2 * slot, key);
      if ((old.lo == 0ull && old.hi == 0ull) || (old.lo == key.lo && old.hi == key.hi)) {
        atomicAdd(counts + slot, 1);
        done = true;
      } else {
        slot = (slot + 1) & mask;
      }
    }
    if (!done) atomicExch(status, 1);
  }
}

__global__ void kmer_collect_kernel(const ulonglong2* __restrict__ table, int64_t capacity, const int32_t* __restrict__ counts,
                                    int k, int min_freq, int64_t* __restrict__ rows, int32_t* __restrict__ freq, int64_t max_out,
                                    unsigned long long* __restrict__ n_out) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < capacity; s += (int64_t)gridDim.x * blockDim.x) {
    const int32_t cnt = counts[s];
    if (cnt < min_freq || cnt <= 0) continue;
    const unsigned long long idx = atomicAdd(n_out, 1ull);
    if ((int64_t)idx >= max_out) continue;                 // the caller sees n_out > max_out and retries with a larger buffer
    const ulonglong2 kv = table[s];
    const unsigned long long m21 = (1ull << 21) - 1ull;
    const unsigned long long v[kKmerMaxK] = {kv.x & m21, (kv.x >> 21) & m21, (kv.x >> 42) & m21,
                                             kv.y & m21, (kv.y >> 21) & m21, (kv.y >> 42) & m21};
#pragma unroll
    for (int j = 0; j < kKmerMaxK; ++j)
      if (j < k) rows[idx * k + j] = (int64_t)v[j];
    freq[idx] = cnt;
  }
}

BinomTable make_binom() {
  BinomTable bt;
  for (int n = 0; n <= kKmerMaxN; ++n)
    for (int j = 0; j <= kKmerMaxK; ++j) {
      if (j == 0) bt.c[n][j] = 1ull;
      else if (n == 0) bt.c[n][j] = 0ull;
      else bt.c[n][j] = bt.c[n - 1][j - 1] + bt.c[n - 1][j];
    }
  return bt;
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

// status word (dev int, zeroed by the caller): 1 = table full, 2 = inconsistent work prefix / cluster above 64 members,
// 3 = node id outside [1, 2^21 - 1]
int matcha_kmer_count(const int64_t* members, const int64_t* offsets, int64_t n_clusters, const int64_t* work_prefix,
                      int64_t total_work, int32_t k, int32_t min_distance, void* table, int64_t capacity, int32_t* counts,
                      int32_t* status, void* stream) {
  MATCHA_REQUIRE(members && offsets && work_prefix && table && counts && status, "kmer_count: NULL argument");
  MATCHA_REQUIRE(k >= 2 && k <= kKmerMaxK, "kmer_count: k=%d unsupported (2..%d)", k, kKmerMaxK);
  MATCHA_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0, "kmer_count: capacity must be a power of two");
  MATCHA_REQUIRE(n_clusters >= 0 && total_work >= 0 && min_distance >= 0, "kmer_count: negative size");
  if (n_clusters == 0 || total_work == 0) return MATCHA_OK;
  static const BinomTable bt = make_binom();
  int64_t blocks = (total_work + 255) / 256;
  if (blocks > kSMs * 32) blocks = kSMs * 32;
  kmer_count_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(members, offsets, n_clusters, work_prefix, k, min_distance, bt,
                                                                  reinterpret_cast<unsigned long long*>(table),
                                                                  (uint64_t)capacity - 1, counts, status);
  MATCHA_CHECK_LAUNCH("kmer_count");
  return MATCHA_OK;
}

// rows dev int64 [max_out, k], freq dev int32 [max_out], n_out dev uint64 [1] (zeroed by the caller; may exceed max_out)
int matcha_kmer_collect(const void* table, int64_t capacity, const int32_t* counts, int32_t k, int32_t min_freq,
                        int64_t* rows, int32_t* freq, int64_t max_out, uint64_t* n_out, void* stream) {
  MATCHA_REQUIRE(table && counts && n_out && (max_out == 0 || (rows && freq)), "kmer_collect: NULL argument");
  MATCHA_REQUIRE(k >= 2 && k <= kKmerMaxK, "kmer_collect: k=%d unsupported (2..%d)", k, kKmerMaxK);
  MATCHA_REQUIRE(capacity > 0, "kmer_collect: empty table");
  int64_t blocks = (capacity + 255) / 256;
  if (blocks > kSMs * 32) blocks = kSMs * 32;
  kmer_collect_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const ulonglong2*>(table), capacity, counts, k,
                                                                    min_freq, rows, freq, max_out,
                                                                    reinterpret_cast<unsigned long long*>(n_out));
  MATCHA_CHECK_LAUNCH("kmer_collect");
  return MATCHA_OK;
}

}  // extern "C"
