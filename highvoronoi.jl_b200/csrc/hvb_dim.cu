// hvb_dim.cu -- one dimension of the search (compiled five times: -DHVB_DIM=2..6), see hvb_ctx.cuh
#include "hvb_ctx.cuh"
#ifndef HVB_DIM
#error "compile with -DHVB_DIM=2..6"
#endif
#define HVB_CAT2(a, b) a##b
#define HVB_CAT(a, b) HVB_CAT2(a, b)
hvb_ctx* HVB_CAT(hvb_make_ctx_, HVB_DIM)() { return new Ctx<HVB_DIM>(); }
