/* forest_synth.c -- generator of the synthetic forest-em corpus (BASELINE configs[4]): every forest
 * gets ITS OWN random AND/OR shape (bench / test tooling, not part of the product library).
 * Shape law (same as carmel_b200/synth.py:_forest_template): root OR; OR fan-out U[2,4] AND children; an AND node is a
 * leaf with probability 0.12 (or at depth 0), else has {1: .3, 2: .6, 3: .1} children; a child is, with probability
 * `share`, a back reference to an earlier multi-node subforest, else an OR (70%) or an AND (30%) one level down.
 * A forest is kept when its hyperedge (AND node) count lies in [0.4, 1.8] x target.
 * Output = the C ABI's cml_forest_batch arrays: next[], label[] (1 for AND, 0 for OR, target index for back references),
 * backref[]; rule ids are assigned by the caller. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t s;
} rng_t;
static inline uint64_t rnext(rng_t* r) {
  uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double runif(rng_t* r) { return (double)(rnext(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rint_(rng_t* r, uint32_t lo, uint32_t hi) { return lo + (uint32_t)(rnext(r) % (hi - lo)); } /* [lo,hi) */

typedef struct {
  uint32_t *nxt, *target;
  uint8_t* kind; /* 0 OR, 1 AND, 2 backref */
  uint32_t n, cap;
  uint32_t* shareable;
  uint32_t n_share, cap_share;
  rng_t* r;
  double share;
  int overflow;
} tmpl_t;

static uint32_t push(tmpl_t* t, uint8_t kind) {
  if (t->n >= t->cap) {
    t->overflow = 1;
    return t->cap - 1;
  }
  t->nxt[t->n] = 0;
  t->kind[t->n] = kind;
  t->target[t->n] = 0;
  return t->n++;
}
static uint32_t node_or(tmpl_t* t, int d);
static uint32_t node_and(tmpl_t* t, int d);
static void child(tmpl_t* t, int d) {
  if (t->n_share && runif(t->r) < t->share) {
    uint32_t i = push(t, 2);
    t->nxt[i] = i + 1;
    t->target[i] = t->shareable[rint_(t->r, 0, t->n_share)];
    return;
  }
  uint32_t i = runif(t->r) < 0.7 ? node_or(t, d) : node_and(t, d);
  if (t->nxt[i] > i + 1 && t->n_share < t->cap_share) t->shareable[t->n_share++] = i;
}
static uint32_t node_and(tmpl_t* t, int d) {
  uint32_t i = push(t, 1);
  if (t->overflow) return i;
  if (d > 0 && runif(t->r) >= 0.12) {
    double u = runif(t->r);
    int k = u < 0.3 ? 1 : (u < 0.9 ? 2 : 3);
    for (int c = 0; c < k && !t->overflow; ++c) child(t, d - 1);
  }
  t->nxt[i] = t->n;
  return i;
}
static uint32_t node_or(tmpl_t* t, int d) {
  uint32_t i = push(t, 0);
  if (t->overflow) return i;
  uint32_t k = rint_(t->r, 2, 5);
  for (uint32_t c = 0; c < k && !t->overflow; ++c) node_and(t, d);
  t->nxt[i] = t->n;
  return i;
}

/* Generates n_forests forests.  Returns the total node count, or 0 when the output capacity (cap_nodes) is too small.
 * node_off[n_forests+1]; next/label/backref[cap_nodes]. */
uint64_t cb200_synth_forests(uint64_t n_forests, uint64_t seed, uint32_t target_he, double share, uint64_t cap_nodes,
                             uint64_t* node_off, uint32_t* next, uint32_t* label, uint8_t* backref) {
  rng_t r = {seed * 0x2545F4914F6CDD1Dull + 0x1234567ull};
  const uint32_t cap = 64 * target_he + 1024;
  tmpl_t t;
  t.nxt = (uint32_t*)malloc(cap * 4);
  t.target = (uint32_t*)malloc(cap * 4);
  t.kind = (uint8_t*)malloc(cap);
  t.shareable = (uint32_t*)malloc(cap * 4);
  t.cap = cap;
  t.cap_share = cap;
  t.r = &r;
  t.share = share;
  uint64_t total = 0;
  node_off[0] = 0;
  for (uint64_t f = 0; f < n_forests;) {
    t.n = 0;
    t.n_share = 0;
    t.overflow = 0;
    node_or(&t, (int)rint_(&r, 4, 8));
    if (t.overflow) continue;
    uint32_t he = 0;
    for (uint32_t i = 0; i < t.n; ++i) he += t.kind[i] == 1;
    if (he < 0.4 * target_he || he > 1.8 * target_he) continue;
    if (total + t.n > cap_nodes) {
      total = 0;
      break;
    }
    for (uint32_t i = 0; i < t.n; ++i) {
      next[total + i] = t.nxt[i];
      backref[total + i] = t.kind[i] == 2;
      label[total + i] = t.kind[i] == 2 ? t.target[i] : (t.kind[i] == 1 ? 1u : 0u);
    }
    total += t.n;
    node_off[++f] = total;
  }
  free(t.nxt);
  free(t.target);
  free(t.kind);
  free(t.shareable);
  return total;
}
