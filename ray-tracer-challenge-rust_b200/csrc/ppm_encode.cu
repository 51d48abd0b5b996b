// ppm_encode.cu — Canvas::to_ppm (canvas.rs:28-58) on the device: the RGBA8 frame the render kernel left in HBM becomes
// the reference's P3 text without the pixels visiting the host first.
//
// The format's only sequential dependence is the 70-column line wrap, and it restarts at every image row
// (canvas.rs:34), so rows are independent:
//   1. ppm_row_bytes_kernel — one WARP per row, 32 channel values at a time: the wrap rule in prefix form (a warp scan + a
//      ballot per line break) counts the row's bytes and notes where every text line starts;
//   2. ppm_row_offsets_kernel — exclusive prefix sum of the row sizes (one block; H is at most a few thousand);
//   3. ppm_line_write_kernel — one CTA per row, one WARP per text LINE (pass 1 recorded where each line starts), one lane
//      per token: digits, single spaces and the line's newline, written in place, neighbouring lanes to neighbouring bytes.
// Byte/integer work, bound by HBM writes of ~12 bytes per pixel; no floating point at all.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "device_scene_impl.cuh"
#include "render.cuh"

namespace rtc {

namespace {

// The wrap rule (canvas.rs:44-55) in prefix form.  A token of n digits costs a = n + 1 characters (its digits and the space
// before it); with P_t the running sum of a over the row's tokens, a line that starts at token s holds the tokens t with
// P_t - P_{s-1} <= 71 (its length is that sum minus the first token's missing space, and the rule allows 70), the byte at
// which it starts inside the row is P_{s-1} (every earlier token's digits and one separator each — space or newline), and
// the row's size with its final newline is P of its last token.
__device__ __forceinline__ unsigned ppm_token_chars(const uchar4* __restrict__ row, unsigned tok, unsigned& value) {
    const uchar4 p = row[tok / 3u];
    const unsigned c = tok % 3u;
    value = c == 0 ? p.x : (c == 1 ? p.y : p.z);
    return 2u + (value >= 10u) + (value >= 100u);
}
__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned v, unsigned lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= (unsigned)off) v += o;
    }
    return v;
}

// pass 1: one WARP per image row, 32 tokens at a time — a warp scan gives every token its P, a ballot finds the first token
// that no longer fits the current line (at most three per 32 tokens: a line holds at least 17), and the lane that owns it
// writes down where its line starts (token index, byte offset inside the row) so that pass 3 can give every line a warp
__global__ void ppm_row_bytes_kernel(const uchar4* __restrict__ px, unsigned width, unsigned height, unsigned max_lines,
                                     unsigned long long* __restrict__ row_bytes, unsigned* __restrict__ row_lines,
                                     uint2* __restrict__ lines) {
    const unsigned y = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (y >= height) return;  // whole warps leave together
    const uchar4* row = px + (size_t)y * width;
    uint2* mine = lines + (size_t)y * max_lines;
    const unsigned ntok = 3u * width;
    unsigned run = 0;   // P of the last token of the chunks before this one
    unsigned base = 0;  // P just before the current line's first token
    unsigned nlines = 1;
    if (lane == 0) mine[0] = make_uint2(0u, 0u);
    for (unsigned t0 = 0; t0 < ntok; t0 += 32) {
        const unsigned t = t0 + lane;
        unsigned v, a = t < ntok ? ppm_token_chars(row, t, v) : 0u;
        const unsigned P = run + warp_inclusive_scan(a, lane);
        run = __shfl_sync(0xffffffffu, P, 31);
        unsigned from = 0;  // lanes >= from are not placed yet
        for (;;) {
            const unsigned over = __ballot_sync(0xffffffffu, t < ntok && lane >= from && P - base > 71u);
            if (!over) break;
            const unsigned j = (unsigned)__ffs((int)over) - 1u;
            base = __shfl_sync(0xffffffffu, P - a, (int)j);  // P of the token before j
            if (lane == j) mine[nlines] = make_uint2(t, base);
            nlines++;
            from = j + 1;  // token j opens the line and always fits
        }
    }
    if (lane == 0) {
        row_bytes[y] = run;  // includes the newline that ends the row (canvas.rs:56)
        row_lines[y] = nlines;
    }
}

// one block: row_off[y] = header + sum of row_bytes[0..y); total at row_off[height]
__global__ void ppm_row_offsets_kernel(const unsigned long long* __restrict__ row_bytes, unsigned height,
                                       unsigned long long header, unsigned long long* __restrict__ row_off) {
    __shared__ unsigned long long part[1024];
    const unsigned t = threadIdx.x, nt = blockDim.x;
    const unsigned per = (height + nt - 1) / nt;
    const unsigned b = t * per, e = b + per < height ? b + per : height;
    unsigned long long s = 0;
    for (unsigned y = b; y < e; y++) s += row_bytes[y];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        unsigned long long acc = header;
        for (unsigned k = 0; k < nt; k++) {
            const unsigned long long v = part[k];
            part[k] = acc;
            acc += v;
        }
        row_off[height] = acc;
    }
    __syncthreads();
    unsigned long long acc = part[t];
    for (unsigned y = b; y < e; y++) {
        row_off[y] = acc;
        acc += row_bytes[y];
    }
}

// pass 3: one CTA per row, one WARP per text line (<= 70 characters, <= 35 tokens), one lane per token: a warp scan places
// every token inside the line, so the lanes' byte stores of one instruction fall side by side in the same few sectors
__global__ void ppm_line_write_kernel(const uchar4* __restrict__ px, unsigned width, unsigned max_lines,
                                      const unsigned long long* __restrict__ row_off,
                                      const unsigned* __restrict__ row_lines, const uint2* __restrict__ lines,
                                      char* __restrict__ out) {
    const unsigned y = blockIdx.x, lane = threadIdx.x & 31u;
    const uchar4* row = px + (size_t)y * width;
    const uint2* mine = lines + (size_t)y * max_lines;
    const unsigned nlines = row_lines[y], ntok = 3u * width;
    char* base = out + row_off[y];
    for (unsigned k = threadIdx.x >> 5; k < nlines; k += blockDim.x >> 5) {
        const uint2 ln = mine[k];
        const unsigned end = (k + 1 < nlines) ? mine[k + 1].x : ntok;
        char* o = base + ln.y;
        unsigned run = 0;  // characters of the line's tokens in the chunks before this one
        for (unsigned t0 = ln.x; t0 < end; t0 += 32) {
            const unsigned t = t0 + lane;
            unsigned v = 0, a = t < end ? ppm_token_chars(row, t, v) : 0u;
            const unsigned p = warp_inclusive_scan(a, lane);
            if (t < end) {
                char* q = o + run + p - a;  // the token's first digit; its separating space sits just before it
                if (t != ln.x) q[-1] = ' ';
                if (v >= 100u) *q++ = (char)('0' + v / 100u);
                if (v >= 10u) *q++ = (char)('0' + (v / 10u) % 10u);
                *q++ = (char)('0' + v % 10u);
                if (t + 1 == end) *q = '\n';  // ends the line — for the row's last line the row terminator (canvas.rs:56)
            }
            run += __shfl_sync(0xffffffffu, p, 31);
        }
    }
}

}  // namespace

uint64_t ppm_max_bytes(uint64_t width, uint64_t height) { return 64 + height * (width * 12 + 1); }

// d_rgba8: width*height uchar4 on `device`.  Writes the PPM text into out_host (capacity bytes; pinned memory makes the
// copy run at PCIe speed) and its length into *len.  If capacity is too small, *len is still set and -1 is returned.
int ppm_encode_device(int device, const void* d_rgba8, uint64_t width, uint64_t height, void* stream_, char* out_host,
                      uint64_t capacity, uint64_t* len, std::string* err) {
    cudaStream_t st = (cudaStream_t)stream_;
    auto fail = [&](const char* what, cudaError_t e) {
        if (err) *err = cuda_err_string(what, e);
        return -3;
    };
    DeviceGuard guard_;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail("cudaSetDevice", e);
    char header[64];
    const int hl = std::snprintf(header, sizeof header, "P3\n%llu %llu\n255\n", (unsigned long long)width,
                                 (unsigned long long)height);
    if (width == 0 || height == 0) {
        *len = (uint64_t)hl;
        if (capacity < *len) return -1;
        std::memcpy(out_host, header, hl);
        return 0;
    }
    unsigned long long *d_bytes = nullptr, *d_off = nullptr;
    unsigned* d_nlines = nullptr;
    uint2* d_lines = nullptr;
    char* d_text = nullptr;
    const uint64_t cap = ppm_max_bytes(width, height);
    // a line holds at least 17 tokens (17 * 4 = 68 <= 70 characters), plus the row's first line
    const unsigned max_lines = (unsigned)((3 * width) / 17 + 2);
    e = cudaMallocAsync((void**)&d_bytes, sizeof(unsigned long long) * height, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_off, sizeof(unsigned long long) * (height + 1), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_nlines, sizeof(unsigned) * height, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_lines, sizeof(uint2) * height * max_lines, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_text, cap, st);
    int rc = 0;
    if (e == cudaSuccess) {
        const unsigned h = (unsigned)height, w = (unsigned)width;
        // four rows (warps) per block
        ppm_row_bytes_kernel<<<(h + 3) / 4, 128, 0, st>>>((const uchar4*)d_rgba8, w, h, max_lines, d_bytes, d_nlines,
                                                          d_lines);
        ppm_row_offsets_kernel<<<1, 1024, 0, st>>>(d_bytes, h, (unsigned long long)hl, d_off);
        ppm_line_write_kernel<<<h, 256, 0, st>>>((const uchar4*)d_rgba8, w, max_lines, d_off, d_nlines, d_lines, d_text);
        e = cudaGetLastError();
    }
    unsigned long long total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_text, header, hl, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_off + height, sizeof(total), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
        *len = total;
        if (total > capacity) rc = -1;
        else {
            e = cudaMemcpyAsync(out_host, d_text, total, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
    }
    if (d_bytes) cudaFreeAsync(d_bytes, st);
    if (d_off) cudaFreeAsync(d_off, st);
    if (d_nlines) cudaFreeAsync(d_nlines, st);
    if (d_lines) cudaFreeAsync(d_lines, st);
    if (d_text) cudaFreeAsync(d_text, st);
    if (e != cudaSuccess) return fail("ppm_encode_device", e);
    return rc;
}

}  // namespace rtc
