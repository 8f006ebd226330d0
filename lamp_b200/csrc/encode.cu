// encode.cu -- order-preserving dictionary coding of the resident table.
//
// The split search only ever ORDERS feature values: min / max over a node (pkg:34-54), `x < cut`
// (pkg:10-32) and the child filters (pkg:1024-1039).  A column with at most 256 distinct values
// (NaN counted as one of them) is therefore stored a second time as one byte per cell.  With
// dict = the column's distinct non-NaN values in ascending order and off = 0 if the column holds a
// NaN, else 1, the stored byte b stands for the "wide code" c = b + off:
//     c == 0      NaN (missing)
//     c == r + 1  the value is dict[r]
// min / max become integer min / max over codes and are decoded through the dictionary, the cutpoint
// min + (max - min) * u is still computed in FP64 from the decoded bounds, and `x < cut` becomes
// `code - 1 < thr` with thr = number of dictionary entries below the cutpoint.  Every decision is
// bit-identical to the FP64 evaluation; the table shrinks 8x (MNIST-shaped 60000 x 784: 376 MB ->
// 47 MB, which stays resident in the 126 MB L2 while the whole forest is built).
//
// -0.0 and +0.0 compare equal in the reference's `<` / `>` and share one dictionary entry (+0.0).
#include "internal.h"

namespace {

constexpr int HT = 1024;  // hash slots per column (the scan stops past 256 distinct values)
constexpr uint64_t EMPTY = 0xffffffffffffffffull;

__device__ __forceinline__ uint32_t hash_slot(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 29;
  return (uint32_t)k & (HT - 1);
}

// One CTA per column: distinct non-NaN values through a shared-memory hash set, then a bitonic
// sort of the (at most 256) survivors.  cnt[col] = number of distinct values, 257 = too many for a byte;
// coff[col] = 0 if the column holds a NaN, else 1.
__global__ void __launch_bounds__(256) k_col_dict(const double *__restrict__ X, int64_t ld, int64_t n,
                                                  double *__restrict__ dict, int32_t *__restrict__ cnt,
                                                  uint8_t *__restrict__ coff) {
  __shared__ unsigned long long s_ht[HT];
  __shared__ double s_val[256];
  __shared__ int s_cnt, s_k, s_nan;
  const int col = blockIdx.x, tid = threadIdx.x;
  const double *c = X + (int64_t)col * ld;
  for (int i = tid; i < HT; i += 256) s_ht[i] = EMPTY;
  if (tid == 0) {
    s_cnt = 0;
    s_k = 0;
    s_nan = 0;
  }
  __syncthreads();
  bool saw_nan = false;
  double prev = __longlong_as_double((long long)EMPTY);  // a NaN: never equal to anything
  for (int64_t r0 = 0; r0 < n; r0 += 256) {
    const int64_t r = r0 + tid;
    if (r < n) {
      const double x = c[r];
      saw_nan |= (x != x);
      if (x == x && !(x == prev)) {
        prev = x;
        const uint64_t key = (uint64_t)__double_as_longlong(x == 0.0 ? 0.0 : x);
        uint32_t h = hash_slot(key);
        for (;;) {
          const unsigned long long cur = s_ht[h];
          if (cur == key) break;
          if (cur == EMPTY) {
            const unsigned long long old = atomicCAS(&s_ht[h], EMPTY, (unsigned long long)key);
            if (old == EMPTY) {
              atomicAdd(&s_cnt, 1);
              break;
            }
            if (old == key) break;
          }
          h = (h + 1) & (HT - 1);
        }
      }
    }
    // at most 256 + 256 slots are ever occupied, so probing always terminates
    if (*(volatile int *)&s_cnt > 256) break;
  }
  if (saw_nan) s_nan = 1;
  __syncthreads();
  const int total = s_cnt;
  const int has_nan = s_nan;
  if (total + has_nan > 256) {
    if (tid == 0) cnt[col] = 257;
    return;
  }
  s_val[tid] = INFINITY;
  __syncthreads();
  for (int i = tid; i < HT; i += 256) {
    const unsigned long long v = s_ht[i];
    if (v != EMPTY) s_val[atomicAdd(&s_k, 1)] = __longlong_as_double((long long)v);
  }
  __syncthreads();
  for (int k = 2; k <= 256; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int partner = tid ^ j;
      if (partner > tid) {
        const double a = s_val[tid], b = s_val[partner];
        const bool up = ((tid & k) == 0);
        if ((a > b) == up) {
          s_val[tid] = b;
          s_val[partner] = a;
        }
      }
      __syncthreads();
    }
  }
  dict[(int64_t)col * 256 + tid] = s_val[tid];  // entries past `total` are +inf
  if (tid == 0) {
    cnt[col] = total;
    coff[col] = has_nan ? 0 : 1;
  }
}

// wide code = 0 for NaN, else 1 + rank of the value in the column's dictionary; stored byte = wide code -
// off.  Each thread codes four consecutive rows: 32-byte reads, one 4-byte write (a warp writes 128
// contiguous bytes).
__global__ void __launch_bounds__(256) k_col_encode(const double *__restrict__ X, int64_t ld, int64_t n,
                                                    const double *__restrict__ dict, const int32_t *__restrict__ cnt,
                                                    const uint8_t *__restrict__ coff, uint8_t *__restrict__ codes,
                                                    int64_t ldc) {
  __shared__ double s_dict[256];
  const int col = blockIdx.y;
  const int total = cnt[col];
  if (total > 256) return;
  const uint32_t off = coff[col];
  s_dict[threadIdx.x] = dict[(int64_t)col * 256 + threadIdx.x];
  __syncthreads();
  const double *c = X + (int64_t)col * ld;
  const int64_t r0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (r0 >= ldc) return;
  uint32_t packed = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int64_t r = r0 + q;
    uint32_t code = 0;
    if (r < n) {
      const double x = c[r];
      if (x == x) {
        int lo = 0, hi = total;  // first entry that is not < x
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_dict[mid] < x)
            lo = mid + 1;
          else
            hi = mid;
        }
        code = (uint32_t)lo + 1u - off;
      }
    }
    packed |= code << (8 * q);
  }
  *reinterpret_cast<uint32_t *>(codes + (int64_t)col * ldc + r0) = packed;
}


// Row-major copy of the byte codes (gathered by k_lane): 32 x 32 tiles through shared memory.
template <typename T>
__global__ void __launch_bounds__(256) k_to_rowmajor(const T *__restrict__ src, int64_t ld_src, int64_t n, int32_t d,
                                                     T *__restrict__ dst, int64_t ld_dst) {
  __shared__ T tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int32_t c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int32_t c = c0 + j;
    const int64_t r = r0 + threadIdx.x;
    if (r < n && c < d) tile[j][threadIdx.x] = src[(int64_t)c * ld_src + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t r = r0 + j;
    const int32_t c = c0 + threadIdx.x;
    if (r < n && c < d) dst[r * ld_dst + c] = tile[threadIdx.x][j];
  }
}

}  // namespace

void et_data_drop_rowmajor(et_data *D) {
  et_dev_free(D->ctx, D->r8, D->r8_bytes);
  D->r8 = nullptr;
  D->r8_bytes = 0;
}

// Builds the row-major copy of the byte codes k_lane gathers from.  A failed allocation just leaves the copy
// absent: the kernels then gather by column.  (A row-major FP64 copy was tried and dropped: with k << d the k values
// of a row sit in as many 32-byte sectors as k column gathers do.)
void et_data_rowmajor(et_ctx *ctx, et_data *D) {
  if (D->n <= 0 || D->d <= 0 || D->coded != 1 || D->r8) return;
  if (const char *e2 = getenv("ETGPU_NO_ROWMAJOR"))
    if (atoi(e2) != 0) return;
  cudaStream_t st = ctx->stream;
  dim3 block(32, 8), grid((unsigned)ceil_div(D->n, 32), (unsigned)ceil_div(D->d, 32));
  D->rs8 = ((int64_t)D->d + 15) / 16 * 16;
  D->r8_bytes = (size_t)D->n * (size_t)D->rs8;
  D->r8 = static_cast<uint8_t *>(et_dev_alloc(ctx, D->r8_bytes));
  if (!D->r8) {
    D->r8_bytes = 0;
    cudaGetLastError();
    return;
  }
  cudaMemsetAsync(D->r8, 0, D->r8_bytes, st);
  k_to_rowmajor<uint8_t><<<grid, block, 0, st>>>(D->c8, D->ldc, D->n, D->d, D->r8, D->rs8);
  ctx->launches++;
}

void et_data_drop_codes(et_data *D) {
  et_data_drop_rowmajor(D);
  et_dev_free(D->ctx, D->c8, (size_t)D->d * (size_t)D->ldc);
  et_dev_free(D->ctx, D->dict, (size_t)D->d * 256 * sizeof(double));
  et_dev_free(D->ctx, D->coff, et_coff_bytes(D->d));
  D->c8 = nullptr;
  D->dict = nullptr;
  D->coff = nullptr;
}

// Builds (or refreshes) the coded copy of the table.  D->coded: 1 = codes valid, -1 = some column has
// more than 256 distinct values (the builder then gathers FP64 values).
void et_data_encode(et_ctx *ctx, et_data *D) {
  if (D->coded != 0) return;
  if (!D->x) {  // CSC table kept sparse: never coded
    D->coded = -1;
    return;
  }
  const char *env = getenv("ETGPU_NO_CODES");
  if ((env && atoi(env) != 0) || D->n <= 0 || D->d <= 0) {
    D->coded = -1;
    et_data_rowmajor(ctx, D);
    return;
  }
  cudaStream_t st = ctx->stream;
  const int64_t n = D->n;
  const int32_t d = D->d;
  D->ldc = ((n + 127) / 128) * 128;
  int32_t *d_cnt = static_cast<int32_t *>(et_dev_alloc(ctx, (size_t)d * sizeof(int32_t)));
  D->dict = static_cast<double *>(et_dev_alloc(ctx, (size_t)d * 256 * sizeof(double)));
  D->coff = static_cast<uint8_t *>(et_dev_alloc(ctx, et_coff_bytes(d)));  // padded with 1 ("no NaN") to whole 16-byte words
  D->c8 = static_cast<uint8_t *>(et_dev_alloc(ctx, (size_t)d * (size_t)D->ldc));
  auto fail = [&](int code, const char *msg) {
    et_dev_free(ctx, d_cnt, (size_t)d * sizeof(int32_t));
    et_data_drop_codes(D);
    cudaGetLastError();
    ET_FAIL(code, "%s", msg);
  };
  if (!d_cnt || !D->dict || !D->coff || !D->c8) fail(ET_ENOMEM, "cannot allocate the coded copy of the table");
  cudaMemsetAsync(D->coff, 1, et_coff_bytes(d), st);
  k_col_dict<<<(unsigned)d, 256, 0, st>>>(D->x, D->ld, n, D->dict, d_cnt, D->coff);
  dim3 grid((unsigned)ceil_div(D->ldc, 1024), (unsigned)d);
  k_col_encode<<<grid, 256, 0, st>>>(D->x, D->ld, n, D->dict, d_cnt, D->coff, D->c8, D->ldc);
  ctx->launches += 2;
  std::vector<int32_t> h_cnt((size_t)d);
  if (cudaMemcpyAsync(h_cnt.data(), d_cnt, (size_t)d * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)
    fail(ET_ECUDA, "coding the table failed");
  et_dev_free(ctx, d_cnt, (size_t)d * sizeof(int32_t));
  d_cnt = nullptr;
  bool ok = true;
  for (int32_t f = 0; f < d; f++) ok &= (h_cnt[(size_t)f] <= 256);
  if (!ok) {
    et_data_drop_codes(D);
    D->coded = -1;
    et_data_rowmajor(ctx, D);
    return;
  }
  D->coded = 1;
  et_data_rowmajor(ctx, D);
}
