// Building blocks shared by the mesh passes (mesh3d.cu: orientation; mt3d.cu: seeded selection):
// an open-addressing table of 16-byte slots keyed by a 64-bit id, and a lock-free union-find.
#pragma once
#include <stdint.h>

namespace ufh {

constexpr unsigned long long EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

struct __align__(16) Slot {                           // key and value in one 16-byte slot: one sector per probe
  unsigned long long key;
  int tri;                                            // payload (mesh3d: smallest triangle on the edge; selection: node index)
  int pad;
};

// slot of `key` in the table (inserting it if `insert`); linear probing, the table always keeps free slots
__device__ __forceinline__ Slot* hash_slot(Slot* tab, size_t mask, unsigned long long key, bool insert) {
  size_t s = (size_t)mix64(key) & mask;
  while (true) {
    unsigned long long k = tab[s].key;
    if (k == key) return tab + s;
    if (k == EMPTY) {
      if (!insert) return tab + s;
      k = atomicCAS(&tab[s].key, EMPTY, key);
      if (k == EMPTY || k == key) return tab + s;
    }
    s = (s + 1) & mask;
  }
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
  while (true) {
    const int p = parent[x];
    if (p == x) return x;
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;                      // path halving (benign race: only ever points nearer the root)
    x = p;
  }
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {                                      // the larger root goes under the smaller: roots are minima
      const int tmp = a;
      a = b;
      b = tmp;
    }
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}

}  // namespace ufh
