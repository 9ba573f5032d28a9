// Candidate record of the nearest-centroid fast pass: the best three (score, index) pairs and the
// fourth-best score of one row over some subset of the codebook.  Ordering is (score descending, index
// ascending), so merging partial records in any order keeps the smallest index among equal scores.
// Three candidates + a certifying fourth score make the exact re-scan a very rare event: the winner can
// only be missed if four fast scores lie within twice the error bound of each other.
#pragma once
#include "common.cuh"

struct Cand {
  float v1, v2, v3, v4;
  int i1, i2, i3;
};

B2T_DEVICE Cand cand_empty() {
  return Cand{-INFINITY, -INFINITY, -INFINITY, -INFINITY, 0x7fffffff, 0x7fffffff, 0x7fffffff};
}

B2T_DEVICE bool cand_better(float v, int i, float w, int j) { return v > w || (v == w && i < j); }

B2T_DEVICE void cand_insert(Cand& c, float v, int idx) {
  if (cand_better(v, idx, c.v1, c.i1)) { c.v4 = c.v3; c.v3 = c.v2; c.i3 = c.i2; c.v2 = c.v1; c.i2 = c.i1; c.v1 = v; c.i1 = idx; }
  else if (cand_better(v, idx, c.v2, c.i2)) { c.v4 = c.v3; c.v3 = c.v2; c.i3 = c.i2; c.v2 = v; c.i2 = idx; }
  else if (cand_better(v, idx, c.v3, c.i3)) { c.v4 = c.v3; c.v3 = v; c.i3 = idx; }
  else if (v > c.v4) { c.v4 = v; }
}

// fast-path insert for strictly increasing indices within one scan (ties keep the earlier index)
B2T_DEVICE void cand_insert_ordered(Cand& c, float v, int idx) {
  if (v > c.v3) {
    if (v > c.v1) { c.v4 = c.v3; c.v3 = c.v2; c.i3 = c.i2; c.v2 = c.v1; c.i2 = c.i1; c.v1 = v; c.i1 = idx; }
    else if (v > c.v2) { c.v4 = c.v3; c.v3 = c.v2; c.i3 = c.i2; c.v2 = v; c.i2 = idx; }
    else { c.v4 = c.v3; c.v3 = v; c.i3 = idx; }
  } else if (v > c.v4) { c.v4 = v; }
}

B2T_DEVICE void cand_merge(Cand& a, const Cand& b) {
  if (b.v1 > -INFINITY) cand_insert(a, b.v1, b.i1);
  if (b.v2 > -INFINITY) cand_insert(a, b.v2, b.i2);
  if (b.v3 > -INFINITY) cand_insert(a, b.v3, b.i3);
  if (b.v4 > a.v4) a.v4 = b.v4;
}
