// synth.cu -- deterministic synthetic read generator for benchmarks and full-size parity tests.
//
// BENCH / TEST TOOLING, not part of the product library.  One counter-based generator, compiled for the
// host and for sm_100a, so that the same bytes can be produced directly in device memory (zero-copy into a
// chunk's planes) or in host memory (for the oracle, the CPU baseline and the end-to-end path).
//
// Model: barcode of read i = whitelist[cells[r]] with r ~ Zipf(s=1) over n_cells with
// probability 0.95, else a uniform whitelist entry (ambient).  Qualities i.i.d. from NovaSeq bins
// {F Q37 .88, ':' Q25 .07, ',' Q11 .04, '#' Q2 .01}.  Each base is substituted with probability
// 10^(-Q/10) by a different base; a '#' base becomes N with probability 0.5.  UMI / prefix bases uniform.
// Record = [prefix_len random bases][barcode][umi_len random bases], optionally with the barcode
// reverse-complemented (neg-strand index reads).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

struct synth_params {
    uint64_t seed;
    uint32_t n_whitelist;
    uint32_t bc_len;
    uint32_t umi_len;
    uint32_t prefix_len;
    uint32_t stride;
    uint32_t n_cells;
    uint32_t reverse;
    uint32_t pad;
};

HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
HD uint64_t rnd(uint64_t seed, uint64_t i, uint64_t j) {
    return mix64(seed + i * 0x9E3779B97F4A7C15ull + (j + 1) * 0xD1B54A32D192ED03ull);
}

// thresholds on a 32-bit uniform for the quality bins, and on a 24-bit uniform for the error rates
HD void qual_and_error(uint64_t r, uint8_t *q, bool *sub, bool *to_n, uint32_t *shift) {
    const uint32_t u = (uint32_t)r;
    const uint32_t e = (uint32_t)(r >> 32) & 0xFFFFFFu;
    uint32_t thr;
    *to_n = false;
    if (u < 3779571220u) { *q = 'F'; thr = 3348u; }              // .88 ; 10^-3.7 * 2^24
    else if (u < 4080218930u) { *q = ':'; thr = 53054u; }        // .07 ; 10^-2.5
    else if (u < 4252017623u) { *q = ','; thr = 1332668u; }      // .04 ; 10^-1.1
    else { *q = '#'; thr = 10585726u; *to_n = ((r >> 56) & 1u) != 0; }   // .01 ; 10^-0.2
    *sub = e < thr;
    *shift = 1u + (uint32_t)((r >> 57) % 3u);
}

HD uint8_t comp_base(uint8_t b) { return b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : b; }

// qual receives the quality bytes [qoff, qoff + qstride) of the record (the plane layout of pcl_chunk_upload)
HD void synth_record(const synth_params &p, const uint8_t *wl, const uint32_t *cells, const uint64_t *zipf_thr, uint64_t i,
                     uint8_t *seq, uint8_t *qual, uint32_t qoff, uint32_t qstride) {
    const char ACGT[4] = {'A', 'C', 'G', 'T'};
    // which barcode
    const uint64_t r0 = rnd(p.seed, i, 0);
    uint32_t widx;
    if ((uint32_t)(r0 >> 32) < 4080218931u) {                     // 0.95: a real cell, Zipf rank by inverse CDF
        const uint64_t u = rnd(p.seed, i, 1);
        uint32_t lo = 0, hi = p.n_cells - 1;
        while (lo < hi) {                                         // first rank with thr[rank] > u
            const uint32_t mid = (lo + hi) >> 1;
            if (zipf_thr[mid] > u) hi = mid; else lo = mid + 1;
        }
        widx = cells[lo];
    } else {
        widx = (uint32_t)(((rnd(p.seed, i, 1) >> 32) * (uint64_t)p.n_whitelist) >> 32);
    }
    const uint8_t *bc = wl + (uint64_t)widx * p.bc_len;
    const uint32_t total = p.prefix_len + p.bc_len + p.umi_len;
    for (uint32_t k = 0; k < total; ++k) {
        const uint64_t r = rnd(p.seed, i, 2 + k);
        uint8_t base;
        if (k >= p.prefix_len && k < p.prefix_len + p.bc_len) {
            const uint32_t j = k - p.prefix_len;
            base = p.reverse ? comp_base(bc[p.bc_len - 1 - j]) : bc[j];
        } else {
            base = (uint8_t)ACGT[(r >> 60) & 3u];
        }
        uint8_t q; bool sub, to_n; uint32_t shift;
        qual_and_error(r, &q, &sub, &to_n, &shift);
        if (to_n) base = 'N';
        else if (sub) {
            const uint32_t c = base == 'A' ? 0u : base == 'C' ? 1u : base == 'G' ? 2u : 3u;
            base = (uint8_t)ACGT[(c + shift) & 3u];
        }
        seq[k] = base;
        if (k >= qoff && k - qoff < qstride) qual[k - qoff] = q;
    }
    for (uint32_t k = total; k < p.stride; ++k) {
        seq[k] = 'N';
        if (k >= qoff && k - qoff < qstride) qual[k - qoff] = '!';
    }
    for (uint32_t k = p.stride > qoff ? p.stride - qoff : 0; k < qstride; ++k) qual[k] = '!';
}

__global__ void k_synth(synth_params p, const uint8_t *wl, const uint32_t *cells, const uint64_t *zipf_thr, uint64_t first,
                        uint64_t n, uint8_t *seq, uint8_t *qual, uint16_t *len, uint32_t qoff, uint32_t qstride) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        synth_record(p, wl, cells, zipf_thr, first + i, seq + i * p.stride, qual + i * (uint64_t)qstride, qoff, qstride);
        len[i] = (uint16_t)(p.prefix_len + p.bc_len + p.umi_len);
    }
}

// sci-RNA-seq3 R1: hairpin barcode (9 or 10 nt) + CAGAGC + UMI 8 + RT barcode 10 (+ 'T' when the hairpin is
// 9 nt, so that every R1 is 34 nt).  cells[] holds (hp_index << 16 | rt_index) pairs.
struct synth_sci3_params {
    uint64_t seed;
    uint32_t n_hp, n_rt, n_cells, stride;
};

HD void synth_sci3_record(const synth_sci3_params &p, const uint8_t *hp, const uint8_t *hp_len, const uint8_t *rt,
                          const uint32_t *cells, const uint64_t *zipf_thr, uint64_t i, uint8_t *seq, uint8_t *qual) {
    const char ACGT[4] = {'A', 'C', 'G', 'T'};
    const uint64_t r0 = rnd(p.seed, i, 0);
    uint32_t hi, ri;
    if ((uint32_t)(r0 >> 32) < 4080218931u) {
        const uint64_t u = rnd(p.seed, i, 1);
        uint32_t lo = 0, top = p.n_cells - 1;
        while (lo < top) {
            const uint32_t mid = (lo + top) >> 1;
            if (zipf_thr[mid] > u) top = mid; else lo = mid + 1;
        }
        hi = cells[lo] >> 16; ri = cells[lo] & 0xFFFFu;
    } else {
        const uint64_t u = rnd(p.seed, i, 1);
        hi = (uint32_t)(((u >> 32) * (uint64_t)p.n_hp) >> 32);
        ri = (uint32_t)(((u & 0xFFFFFFFFull) * (uint64_t)p.n_rt) >> 32);
    }
    const uint32_t hl = hp_len[hi];
    const char *linker = "CAGAGC";
    for (uint32_t k = 0; k < 34u; ++k) {
        const uint64_t r = rnd(p.seed, i, 2 + k);
        uint8_t base;
        if (k < hl) base = hp[hi * 10u + k];
        else if (k < hl + 6u) base = (uint8_t)linker[k - hl];
        else if (k < hl + 14u) base = (uint8_t)ACGT[(r >> 60) & 3u];
        else if (k < hl + 24u) base = rt[ri * 10u + (k - hl - 14u)];
        else base = 'T';
        uint8_t q; bool sub, to_n; uint32_t shift;
        qual_and_error(r, &q, &sub, &to_n, &shift);
        if (to_n) base = 'N';
        else if (sub) {
            const uint32_t c = base == 'A' ? 0u : base == 'C' ? 1u : base == 'G' ? 2u : 3u;
            base = (uint8_t)ACGT[(c + shift) & 3u];
        }
        seq[k] = base;
        qual[k] = q;
    }
    for (uint32_t k = 34u; k < p.stride; ++k) { seq[k] = 'N'; qual[k] = '!'; }
}

__global__ void k_synth_sci3(synth_sci3_params p, const uint8_t *hp, const uint8_t *hp_len, const uint8_t *rt,
                             const uint32_t *cells, const uint64_t *zipf_thr, uint64_t first, uint64_t n, uint8_t *seq,
                             uint8_t *qual, uint16_t *len) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        synth_sci3_record(p, hp, hp_len, rt, cells, zipf_thr, first + i, seq + i * p.stride, qual + i * p.stride);
        len[i] = 34;
    }
}

__global__ void k_fill_u16(uint16_t *dst, uint16_t v, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = v;
}

struct synth_t {
    synth_params p;
    std::vector<uint8_t> wl;
    std::vector<uint32_t> cells;
    std::vector<uint64_t> thr;
    uint8_t *d_wl = nullptr;
    uint32_t *d_cells = nullptr;
    uint64_t *d_thr = nullptr;
};

extern "C" {

synth_t *synth_create(const synth_params *p, const uint8_t *wl_bytes, const uint32_t *cells, const uint64_t *zipf_thr) {
    synth_t *s = new synth_t();
    s->p = *p;
    s->wl.assign(wl_bytes, wl_bytes + (uint64_t)p->n_whitelist * p->bc_len);
    s->cells.assign(cells, cells + p->n_cells);
    s->thr.assign(zipf_thr, zipf_thr + p->n_cells);
    return s;
}

void synth_destroy(synth_t *s) {
    if (!s) return;
    if (s->d_wl) cudaFree(s->d_wl);
    if (s->d_cells) cudaFree(s->d_cells);
    if (s->d_thr) cudaFree(s->d_thr);
    delete s;
}

int synth_fill_host(synth_t *s, uint64_t first, uint64_t n, uint8_t *seq, uint8_t *qual, uint16_t *len, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int t) {
        uint64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
        for (uint64_t i = lo; i < hi; ++i) {
            synth_record(s->p, s->wl.data(), s->cells.data(), s->thr.data(), first + i, seq + i * s->p.stride, qual + i * s->p.stride,
                         0, s->p.stride);
            len[i] = (uint16_t)(s->p.prefix_len + s->p.bc_len + s->p.umi_len);
        }
    };
    if (n_threads == 1) { work(0); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto &t : th) t.join();
    return 0;
}

// device pointers; stream is a cudaStream_t.  Returns 0 or a cudaError_t.
int synth_fill_device(synth_t *s, int device, uint64_t first, uint64_t n, void *seq, void *qual, void *len, void *stream,
                      uint32_t qoff, uint32_t qstride) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    if (!s->d_wl) {
        if ((e = cudaMalloc(&s->d_wl, s->wl.size())) != cudaSuccess) return (int)e;
        if ((e = cudaMalloc(&s->d_cells, s->cells.size() * 4)) != cudaSuccess) return (int)e;
        if ((e = cudaMalloc(&s->d_thr, s->thr.size() * 8)) != cudaSuccess) return (int)e;
        cudaMemcpy(s->d_wl, s->wl.data(), s->wl.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(s->d_cells, s->cells.data(), s->cells.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(s->d_thr, s->thr.data(), s->thr.size() * 8, cudaMemcpyHostToDevice);
    }
    if (n == 0) return 0;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_synth<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(s->p, s->d_wl, s->d_cells, s->d_thr, first, n, (uint8_t *)seq,
                                                               (uint8_t *)qual, (uint16_t *)len, qoff, qstride);
    return (int)cudaGetLastError();
}

struct synth3_t {
    synth_sci3_params p;
    std::vector<uint8_t> hp, hp_len, rt;
    std::vector<uint32_t> cells;
    std::vector<uint64_t> thr;
    uint8_t *d_hp = nullptr, *d_hl = nullptr, *d_rt = nullptr;
    uint32_t *d_cells = nullptr;
    uint64_t *d_thr = nullptr;
};

synth3_t *synth3_create(const synth_sci3_params *p, const uint8_t *hp10, const uint8_t *hp_len, const uint8_t *rt10,
                        const uint32_t *cells, const uint64_t *zipf_thr) {
    synth3_t *s = new synth3_t();
    s->p = *p;
    s->hp.assign(hp10, hp10 + (size_t)p->n_hp * 10);
    s->hp_len.assign(hp_len, hp_len + p->n_hp);
    s->rt.assign(rt10, rt10 + (size_t)p->n_rt * 10);
    s->cells.assign(cells, cells + p->n_cells);
    s->thr.assign(zipf_thr, zipf_thr + p->n_cells);
    return s;
}

void synth3_destroy(synth3_t *s) {
    if (!s) return;
    cudaFree(s->d_hp); cudaFree(s->d_hl); cudaFree(s->d_rt); cudaFree(s->d_cells); cudaFree(s->d_thr);
    delete s;
}

int synth3_fill_host(synth3_t *s, uint64_t first, uint64_t n, uint8_t *seq, uint8_t *qual, uint16_t *len, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int t) {
        uint64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
        for (uint64_t i = lo; i < hi; ++i) {
            synth_sci3_record(s->p, s->hp.data(), s->hp_len.data(), s->rt.data(), s->cells.data(), s->thr.data(), first + i,
                              seq + i * s->p.stride, qual + i * s->p.stride);
            len[i] = 34;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto &t : th) t.join();
    return 0;
}

int synth3_fill_device(synth3_t *s, int device, uint64_t first, uint64_t n, void *seq, void *qual, void *len, void *stream) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    if (!s->d_hp) {
        cudaMalloc(&s->d_hp, s->hp.size()); cudaMalloc(&s->d_hl, s->hp_len.size()); cudaMalloc(&s->d_rt, s->rt.size());
        cudaMalloc(&s->d_cells, s->cells.size() * 4); cudaMalloc(&s->d_thr, s->thr.size() * 8);
        cudaMemcpy(s->d_hp, s->hp.data(), s->hp.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(s->d_hl, s->hp_len.data(), s->hp_len.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(s->d_rt, s->rt.data(), s->rt.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(s->d_cells, s->cells.data(), s->cells.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(s->d_thr, s->thr.data(), s->thr.size() * 8, cudaMemcpyHostToDevice);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    }
    if (n == 0) return 0;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_synth_sci3<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(s->p, s->d_hp, s->d_hl, s->d_rt, s->d_cells, s->d_thr, first, n,
                                                                    (uint8_t *)seq, (uint8_t *)qual, (uint16_t *)len);
    return (int)cudaGetLastError();
}

int synth_memcpy_d2h(int device, void *dst, const void *src, uint64_t bytes) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
}

int synth_fill_u16_device(int device, void *dst, uint16_t v, uint64_t n, void *stream) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return 0;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_fill_u16<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint16_t *)dst, v, n);
    return (int)cudaGetLastError();
}

}  // extern "C"
