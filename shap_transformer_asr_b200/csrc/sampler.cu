// Host-side (native, no device code): the random half of the KernelSHAP coalition sampler.
//
// shap.KernelExplainer draws its random coalitions from the GLOBAL legacy numpy generator -- one
// np.random.permutation(M) per draw -- so "the same coalition index sets for the same
// seed" means reproducing numpy's MT19937 stream and its legacy shuffle exactly.  The Python loop that did this cost
// 9.5 ms per 2048-coalition clip (47 ms for 8192), every rank paying it in full: at 8 GPUs it was a third of the
// seconds-per-explained-clip.  Here the same integer arithmetic runs natively on the generator state handed over by
// np.random.get_state() and handed back through np.random.set_state(); all floating-point parts of the sampler
// (size weights, np.random.choice, the final weight scaling) stay in numpy, so weights are bit-identical by construction.
//
//   mt19937 output + tempering ........ numpy/random/src/mt19937/mt19937.h (mt19937_next), standard MT19937
//   bounded integer ................... numpy/random/src/distributions/distributions.c: random_interval (mask + reject)
//   permutation ....................... RandomState.permutation(int) = arange + RandomState._shuffle_raw:
//                                       for i in reversed(range(1, n)): j = random_interval(i); swap(x[i], x[j])
#include "../../include/w2s.h"

#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Mt19937 {
  uint32_t* key;   // [624], caller-owned
  int pos;
  void regenerate() {
    constexpr int N = 624, M = 397;
    constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    uint32_t y;
    int kk;
    for (kk = 0; kk < N - M; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; kk < N - 1; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    pos = 0;
  }
  inline uint32_t next() {
    if (pos >= 624) regenerate();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // legacy random_interval for max <= 0xffffffff: smallest all-ones mask >= max, reject values above max
  inline uint32_t interval(uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    while ((v = (next() & mask)) > max) {
    }
    return v;
  }
};

}  // namespace

extern "C" int64_t w2s_sample_rows(int M, int n_full, int n_paired, const int64_t* ind_set, int64_t n_ind,
                                   int64_t samples_left, int64_t added, uint32_t* mt_key, int32_t* mt_pos,
                                   uint32_t* z_words, double* weights, int64_t cap_rows, int64_t* ind_used) {
  if (M < 2 || M > 2048 || !ind_set || !mt_key || !mt_pos || !z_words || !weights) return -1;
  const int W = (M + 31) / 32;
  Mt19937 rng{mt_key, *mt_pos};
  std::vector<int32_t> perm((size_t)M);
  std::vector<uint32_t> row((size_t)W);
  std::unordered_map<std::string, int64_t> seen;
  seen.reserve((size_t)(samples_left * 2 + 16));
  const uint32_t tail = (M % 32) ? ((1u << (M % 32)) - 1u) : 0xffffffffu;   // valid bits of the last word
  int64_t pos = 0;
  while (samples_left > 0 && pos < n_ind) {
    const int size = (int)ind_set[pos] + n_full + 1;
    ++pos;
    if (size < 1 || size > M) return -1;
    for (int i = 0; i < M; ++i) perm[(size_t)i] = i;
    for (int i = M - 1; i >= 1; --i) {
      const uint32_t j = rng.interval((uint32_t)i);
      const int32_t t = perm[j];
      perm[j] = perm[(size_t)i];
      perm[(size_t)i] = t;
    }
    std::fill(row.begin(), row.end(), 0u);
    for (int i = 0; i < size; ++i) row[(size_t)(perm[(size_t)i] >> 5)] |= 1u << (perm[(size_t)i] & 31);
    std::string key(reinterpret_cast<const char*>(row.data()), (size_t)W * 4);
    auto it = seen.find(key);
    const bool fresh = it == seen.end();
    int64_t at = 0;
    if (fresh) {
      if (added >= cap_rows) return -1;
      seen.emplace(std::move(key), added);
      --samples_left;
      std::memcpy(z_words + added * W, row.data(), (size_t)W * 4);
      weights[added] = 1.0;
      ++added;
    } else {
      at = it->second;
      weights[at] += 1.0;
    }
    if (samples_left > 0 && size <= n_paired) {
      if (fresh) {
        if (added >= cap_rows) return -1;
        --samples_left;
        uint32_t* dst = z_words + added * W;
        for (int w = 0; w < W; ++w) dst[w] = ~row[(size_t)w];
        dst[W - 1] &= tail;
        weights[added] = 1.0;
        ++added;
      } else {
        weights[at + 1] += 1.0;
      }
    }
  }
  *mt_pos = rng.pos;
  if (ind_used) *ind_used = pos;
  return added;
}
