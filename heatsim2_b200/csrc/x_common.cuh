// x_common.cuh - pieces shared by the x-sweep kernels (kernels_xf.cu).
#pragma once
#include "chunk_core.cuh"

struct SrcTab {  // active volumetric classes of this step (<= 8)
  int n;
  uint8_t idx[8];
  double val[8];
};

inline int hs2_make_src_tab(const hs2_source *src, SrcTab *st) {
  st->n = 0;
  if (!src || !src->h_value || !src->d_vol_elements) return HS2_OK;
  for (int v = 0; v < 256; ++v) {
    if (src->h_value[v] != 0.0) {
      HS2_REQUIRE(st->n < 8, "hs2_source: more than 8 active volumetric classes in one step; pass a dense source array instead");
      st->idx[st->n] = (uint8_t)v;
      st->val[st->n] = src->h_value[v];
      st->n++;
    }
  }
  return HS2_OK;
}

