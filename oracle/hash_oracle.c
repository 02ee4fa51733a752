/* CPU ORACLE (C restatement) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Sequential restatement of the reference's hash-table / LFU-cache integer algorithms, used by
 * tests/ to check large key sets quickly (the numpy oracle loops in Python) and as a second,
 * independent implementation beside oracle/tt_oracle.py.  Built by __graft_entry__.build() with gcc into
 * oracle/_build/libhash_oracle.so.  The product never links or loads it.
 *
 *   ttb_oracle_hash            hashtbl_cuda_utils.cuh:48-76   murmur3 (lo word, hi word, h ^= 2, fmix32) + Lemire range
 *   ttb_oracle_find            hashtbl_cuda_utils.cuh:135-154 linear probe <= 3, never stops at an empty slot (Q2)
 *   ttb_oracle_update          tt_embeddings_cuda.cu:1077-1089 + hashtbl_cuda_utils.cuh:102-133, in index order
 *   ttb_oracle_mark_popular    tt_embeddings_cuda.cu:1115-1139 applied to a stably frequency-sorted slot list
 */
#include <stdint.h>
#include <stdlib.h>

#define MAX_PROBES 3
#define UNUSED_KEY (-1LL)

static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t ttb_oracle_hash(int64_t key, int32_t C) {
  const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  uint32_t h = 0;
  uint32_t w[2] = {(uint32_t)((uint64_t)key & 0xffffffffu), (uint32_t)((uint64_t)key >> 32)};
  for (int i = 0; i < 2; ++i) {
    uint32_t k = w[i];
    k *= c1;
    k = rotl32(k, 15);
    k *= c2;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5 + 0xe6546b64u;
  }
  h ^= 2;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return (uint32_t)(((uint64_t)h * (uint64_t)(uint32_t)C) >> 32);
}

int32_t ttb_oracle_find(int64_t key, const int64_t* tbl, int32_t C) {
  int32_t slot = (int32_t)ttb_oracle_hash(key, C);
  for (int n = 0; n < MAX_PROBES; ++n) {
    if (tbl[slot] == key) return slot;
    if (key == UNUSED_KEY) return -1;
    slot = (slot + 1) % C;
  }
  return -1;
}

/* returns the number of dropped (not inserted) lookups */
int64_t ttb_oracle_update(const int64_t* idx, int64_t nnz, int64_t* tbl, int64_t* freq, int32_t C) {
  int64_t dropped = 0;
  for (int64_t i = 0; i < nnz; ++i) {
    const int64_t key = idx[i];
    int32_t slot = (int32_t)ttb_oracle_hash(key, C);
    int placed = 0;
    for (int n = 0; n < MAX_PROBES && !placed; ++n) {
      int64_t old = tbl[slot];
      if (old == UNUSED_KEY) {
        tbl[slot] = key;
        old = key;
      }
      if (old == key) {
        freq[slot] += 1;
        placed = 1;
      } else {
        slot = (slot + 1) % C;
      }
    }
    dropped += !placed;
  }
  return dropped;
}

typedef struct {
  int64_t freq;
  int32_t slot;
} slot_freq_t;

static int cmp_desc_stable(const void* a, const void* b) {
  const slot_freq_t* x = (const slot_freq_t*)a;
  const slot_freq_t* y = (const slot_freq_t*)b;
  if (x->freq != y->freq) return x->freq > y->freq ? -1 : 1;
  return x->slot < y->slot ? -1 : (x->slot > y->slot);  /* stable: ties keep slot order */
}

/* K12 + K13: sorted_keys[n] = key of the n-th most frequent slot; marks / evicts in place */
void ttb_oracle_populate_state(int64_t cache_size, int64_t* tbl, int64_t* freq, int32_t* state, int32_t C,
                               int64_t* sorted_keys) {
  slot_freq_t* order = (slot_freq_t*)malloc(sizeof(slot_freq_t) * (size_t)C);
  for (int32_t s = 0; s < C; ++s) {
    order[s].freq = freq[s];
    order[s].slot = s;
  }
  qsort(order, (size_t)C, sizeof(slot_freq_t), cmp_desc_stable);
  int64_t* snapshot = (int64_t*)malloc(sizeof(int64_t) * (size_t)C);
  for (int32_t n = 0; n < C; ++n) snapshot[n] = sorted_keys[n] = tbl[order[n].slot];
  for (int32_t n = 0; n < C; ++n) {
    const int64_t key = snapshot[n];
    if (key != UNUSED_KEY) {
      const int32_t slot = ttb_oracle_find(key, tbl, C);
      if (slot < 0) continue;
      if (n < cache_size) {
        state[slot] = n;
      } else {
        tbl[slot] = UNUSED_KEY;
        freq[slot] = 0;
      }
    } else if (n < cache_size) {
      sorted_keys[n] = 0;
    }
  }
  free(snapshot);
  free(order);
}
