// cuda.h -- CPU EMULATION of the driver-API slice used by csrc/fs2d_fused.cu / fs2d_stream.cu (tensor maps).
// TEST INFRASTRUCTURE ONLY, see cuda_runtime.h.
#pragma once
#include "cuda_runtime.h"

typedef int CUresult;
constexpr CUresult CUDA_SUCCESS = 0;
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

struct alignas(64) CUtensorMap {       // the real one is an opaque 128-byte descriptor
    const char *base;
    uint64_t dims[2];                  // elements: {innermost (columns), rows}
    uint64_t row_stride;               // bytes
    uint32_t box[2];                   // elements: {columns, rows}
    uint32_t esz;
    uint32_t pad[19];
};
static_assert(sizeof(CUtensorMap) == 128, "descriptor size");

namespace emu {
// the hardware's own requirements, so that a descriptor the driver would refuse is refused here too
inline CUresult encode_tiled(CUtensorMap *m, CUtensorMapDataType dt, cuuint32_t rank, void *base, const cuuint64_t *dims,
                             const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *estr, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    const uint32_t esz = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 1;
    if (rank != 2 || estr[0] != 1 || estr[1] != 1) return 1;
    if ((uintptr_t)base % 16 != 0 || strides[0] % 16 != 0) return 1;            // global address / stride alignment
    if (box[0] == 0 || box[1] == 0 || box[0] > 256 || box[1] > 256) return 1;   // box dimensions <= 256
    if ((box[0] * esz) % 16 != 0) return 1;                                     // inner box extent: multiple of 16 bytes
    m->base = (const char *)base;
    m->dims[0] = dims[0];
    m->dims[1] = dims[1];
    m->row_stride = strides[0];
    m->box[0] = box[0];
    m->box[1] = box[1];
    m->esz = esz;
    return CUDA_SUCCESS;
}
// cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes: box at element coordinates (c0, c1),
// out-of-bounds elements are zero-filled, the whole box counts towards the transaction bytes
inline void tma_load_2d(void *dst, const CUtensorMap *m, int c0, int c1, uint64_t *bar) {
    if (((long long)c0 * m->esz) % 16 != 0) {   // measured on B200 (scripts/probes/tma_probe.cu): illegal instruction
        fprintf(stderr, "cuda_emu: TMA box start %d x %u bytes is not 16-byte aligned in the innermost dimension\n", c0, m->esz);
        abort();
    }
    if ((uintptr_t)dst % 128 != 0) {
        fprintf(stderr, "cuda_emu: TMA shared-memory destination is not 128-byte aligned\n");
        abort();
    }
    char *d = (char *)dst;
    for (uint32_t r = 0; r < m->box[1]; ++r)
        for (uint32_t c = 0; c < m->box[0]; ++c) {
            const long long gc = (long long)c0 + c, gr = (long long)c1 + r;
            char *o = d + ((size_t)r * m->box[0] + c) * m->esz;
            if (gc >= 0 && gr >= 0 && gc < (long long)m->dims[0] && gr < (long long)m->dims[1])
                memcpy(o, m->base + (size_t)gr * m->row_stride + (size_t)gc * m->esz, m->esz);
            else
                memset(o, 0, m->esz);
        }
    MBar *b = reinterpret_cast<MBar *>(bar);
    b->tx -= m->box[0] * m->box[1] * m->esz;
    mbar_check(b);
}
}  // namespace emu

inline cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, int, cudaDriverEntryPointQueryResult *q) {
    *fn = strcmp(name, "cuTensorMapEncodeTiled") == 0 ? (void *)emu::encode_tiled : nullptr;
    *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}
