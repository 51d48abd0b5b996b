// ppm_encode.cu — Canvas::to_ppm (canvas.rs:28-58) on the device: the RGBA8 frame the render kernel left in HBM becomes
// the reference's P3 text without the pixels visiting the host first.
//
// The format's only sequential dependence is the 70-column line wrap, and it restarts at every image row
// (canvas.rs:34), so rows are independent:
//   1. ppm_row_bytes_kernel — one thread per row walks the row's 3*W channel values with the reference's wrap rule,
//      counts the row's bytes and notes where every text line starts;
//   2. ppm_row_offsets_kernel — exclusive prefix sum of the row sizes (one block; H is at most a few thousand);
//   3. ppm_line_write_kernel — one CTA per row, one thread per text LINE (pass 1 recorded where each line starts):
//      digits, single spaces and the line's newline, written in place.
// Byte/integer work, bound by HBM writes of ~12 bytes per pixel; no floating point at all.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "device_scene_impl.cuh"
#include "render.cuh"

namespace rtc {

namespace {

// canvas.rs:44-55 for one token of n digits: returns the bytes it adds (separator + digits) and updates the line length
__device__ __forceinline__ unsigned ppm_token(unsigned& len, unsigned n, bool& newline) {
    unsigned add = n;
    newline = false;
    if (len + n + 1 > 70) {
        newline = true;
        add += 1;
        len = 0;
    }
    if (len > 0) {
        add += 1;  // the separating space (a newline resets len to 0, so the two never combine)
        len += 1;
    }
    len += n;
    return add;
}

// pass 1: the wrap rule is a sequential walk of the row's tokens — do it once, per row, and write down where every LINE of
// the row starts (token index and byte offset inside the row) so that pass 3 can give every line its own thread
__global__ void ppm_row_bytes_kernel(const uchar4* __restrict__ px, unsigned width, unsigned height, unsigned max_lines,
                                     unsigned long long* __restrict__ row_bytes, unsigned* __restrict__ row_lines,
                                     uint2* __restrict__ lines) {
    const unsigned y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= height) return;
    const uchar4* row = px + (size_t)y * width;
    uint2* mine = lines + (size_t)y * max_lines;
    unsigned len = 0, bytes = 0, nlines = 0, tok = 0;
    mine[nlines++] = make_uint2(0u, 0u);
    for (unsigned x = 0; x < width; x++) {
        const uchar4 p = row[x];
        const unsigned v[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int c = 0; c < 3; c++, tok++) {
            const unsigned n = 1u + (v[c] >= 10u) + (v[c] >= 100u);
            bool nl;
            const unsigned add = ppm_token(len, n, nl);
            if (nl) mine[nlines++] = make_uint2(tok, bytes + 1);  // the line starts after the newline character
            bytes += add;
        }
    }
    row_bytes[y] = (unsigned long long)bytes + 1;  // the newline that ends the row (canvas.rs:56)
    row_lines[y] = nlines;
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

// pass 3: one CTA per row, one thread per text line (<= 70 characters): digits, single spaces, the line's newline
__global__ void ppm_line_write_kernel(const uchar4* __restrict__ px, unsigned width, unsigned max_lines,
                                      const unsigned long long* __restrict__ row_off,
                                      const unsigned* __restrict__ row_lines, const uint2* __restrict__ lines,
                                      char* __restrict__ out) {
    const unsigned y = blockIdx.x;
    const uchar4* row = px + (size_t)y * width;
    const uint2* mine = lines + (size_t)y * max_lines;
    const unsigned nlines = row_lines[y], ntok = 3u * width;
    char* base = out + row_off[y];
    for (unsigned k = threadIdx.x; k < nlines; k += blockDim.x) {
        const uint2 ln = mine[k];
        const unsigned end = (k + 1 < nlines) ? mine[k + 1].x : ntok;
        char* o = base + ln.y;
        for (unsigned tok = ln.x; tok < end; tok++) {
            const uchar4 p = row[tok / 3u];
            const unsigned c = tok % 3u;
            const unsigned v = c == 0 ? p.x : (c == 1 ? p.y : p.z);
            if (tok != ln.x) *o++ = ' ';
            if (v >= 100u) *o++ = (char)('0' + v / 100u);
            if (v >= 10u) *o++ = (char)('0' + (v / 10u) % 10u);
            *o++ = (char)('0' + v % 10u);
        }
        *o = '\n';  // ends the line — for the row's last line this is the row terminator (canvas.rs:56)
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
        // 32-thread blocks: rows are long sequential walks, spread them over as many SMs as there are
        ppm_row_bytes_kernel<<<(h + 31) / 32, 32, 0, st>>>((const uchar4*)d_rgba8, w, h, max_lines, d_bytes, d_nlines,
                                                           d_lines);
        ppm_row_offsets_kernel<<<1, 1024, 0, st>>>(d_bytes, h, (unsigned long long)hl, d_off);
        ppm_line_write_kernel<<<h, 128, 0, st>>>((const uchar4*)d_rgba8, w, max_lines, d_off, d_nlines, d_lines, d_text);
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
