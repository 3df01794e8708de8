// nccl_dyn.h -- NCCL is loaded with dlopen at run time (libnccl.so.2), only when a multi-GPU context is created.
// The single-GPU library therefore has no link-time NCCL dependency, and inside a torchrun process the NCCL that
// torch already loaded is reused.
#pragma once
#include <stddef.h>
#include <stdint.h>

struct cniic_ctx;

int cniic_nccl_init(cniic_ctx *ctx, int rank, int world, const uint8_t unique_id[128]);
void cniic_nccl_destroy(cniic_ctx *ctx);
// in-place sum all-reduce of `count` u64 values on the ctx stream
int cniic_nccl_allreduce_u64(cniic_ctx *ctx, unsigned long long *d_buf, size_t count);
int cniic_nccl_allgather_bytes(cniic_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank);
