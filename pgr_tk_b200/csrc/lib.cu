// lib.cu — single translation unit of libpgr_b200.so (kernels are defined in headers shared by the parts below).
#include "ctx.cu"
#include "index.cu"
#include "query.cu"
#include "bundles.cu"
#include "frags.cu"
#include "shard.cu"
