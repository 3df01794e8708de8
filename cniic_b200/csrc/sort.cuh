#pragma once
#include "common.cuh"
// Sorts the n RGB points by the Morton code of their colour: d_sorted[i] = packed colour (r | g<<8 | b<<16) of the i-th
// point in sorted order, d_perm[i] = its original index, d_wsorted[i] = its weight (if d_wts != nullptr).
int cniic_dev_sort_colours(cniic_ctx *ctx, const uint8_t *d_rgb, const uint32_t *d_wts, size_t n, uint32_t *d_sorted, uint32_t *d_perm,
                           uint32_t *d_wsorted, uint32_t *launches);
