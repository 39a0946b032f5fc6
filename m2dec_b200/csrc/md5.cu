// md5.cu — per-picture MD5 of the cropped NV12 output on the device.
//
// What the reference application hashes for `h264dec -O` (src/app/filewrite.h:11-29, 91-124 with
// src/app/md5.c): the cropped luma rows top to bottom, then the cropped rows of the interleaved
// CbCr plane, one digest per output frame.  MD5 (RFC 1321) is a serial chain over 64-byte blocks,
// so the parallelism is across pictures: one thread per picture, thousands of pictures per
// launch (kernel replay keeps them resident).  That is slow per picture (tens of milliseconds)
// and fast per launch: 16 bytes per picture cross PCIe instead of 3 MB.
#include <stdint.h>
#include "device_common.cuh"

namespace b2 {
namespace {

__constant__ uint32_t c_md5_k[64] = {
	0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be,
	0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
	0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
	0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
	0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
	0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};

__device__ __forceinline__ uint32_t rotl(uint32_t x, int s) { return __funnelshift_l(x, x, s); }

__device__ void md5_block(uint32_t st[4], const uint32_t m[16]) {
	uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#pragma unroll
	for (int i = 0; i < 64; ++i) {
		uint32_t f;
		int g, s;
		if (i < 16) { f = (b & c) | (~b & d); g = i; s = (i & 3) == 0 ? 7 : (i & 3) == 1 ? 12 : (i & 3) == 2 ? 17 : 22; }
		else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; s = (i & 3) == 0 ? 5 : (i & 3) == 1 ? 9 : (i & 3) == 2 ? 14 : 20; }
		else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; s = (i & 3) == 0 ? 4 : (i & 3) == 1 ? 11 : (i & 3) == 2 ? 16 : 23; }
		else { f = c ^ (b | ~d); g = (7 * i) & 15; s = (i & 3) == 0 ? 6 : (i & 3) == 1 ? 10 : (i & 3) == 2 ? 15 : 21; }
		const uint32_t t = a + f + c_md5_k[i] + m[g];
		a = d; d = c; c = b;
		b = b + rotl(t, s);
	}
	st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

// The message as a byte stream over the two cropped planes.  `stride` is the row pitch the APPLICATION reads the planes
// with (m2d_frame_t.width: the reference's writer, filewrite.h:99-124, addresses both planes with it), `plane` the
// distance from the luma plane to the CbCr plane.
struct Cursor {
	const uint8_t *luma, *chroma;
	int stride, left, top, w, h;     // window in luma samples; chroma rows: top / 2 .. top / 2 + h / 2 - 1, same byte columns
	int plane, row, col;
	__device__ const uint8_t *ptr() const {
		return plane == 0 ? luma + (size_t)(top + row) * stride + left + col : chroma + (size_t)((top >> 1) + row) * stride + left + col;
	}
	__device__ void advance(int n) {      // n bytes inside the current row
		col += n;
		if (col >= w) {
			col = 0;
			if (++row >= (plane == 0 ? h : h >> 1)) { row = 0; ++plane; }
		}
	}
};

__global__ void md5_pictures_kernel(const uint8_t *__restrict__ frames, size_t frame_bytes, const int *__restrict__ slots, int slot0, int n, int stride, size_t plane_bytes,
                                    int left, int top, int w, int h, uint8_t *__restrict__ out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Cursor cu;
	cu.luma = frames + frame_bytes * (size_t)(slots ? slots[i] : slot0 + i);
	cu.chroma = cu.luma + plane_bytes;
	cu.stride = stride; cu.left = left; cu.top = top; cu.w = w; cu.h = h;
	cu.plane = 0; cu.row = 0; cu.col = 0;
	uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
	const uint64_t total = (uint64_t)w * h + (uint64_t)w * (h >> 1);
	// every 64-byte block lies inside one row, 16-byte aligned: four 128-bit loads per block, and the next block is
	// asked for before the current one is hashed (the chain of 64 dependent steps hides the load)
	const bool aligned = (w & 63) == 0 && ((left | stride) & 15) == 0 && (plane_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(cu.luma) & 15) == 0;
	uint64_t done = 0;
	uint32_t m[16];
	if (aligned && total >= 64) {
		uint4 nx[4];
		{
			const uint4 *p = reinterpret_cast<const uint4 *>(cu.ptr());
#pragma unroll
			for (int k = 0; k < 4; ++k) nx[k] = __ldg(p + k);
			cu.advance(64);
		}
		for (; done + 64 <= total; done += 64) {
#pragma unroll
			for (int k = 0; k < 4; ++k) { m[4 * k] = nx[k].x; m[4 * k + 1] = nx[k].y; m[4 * k + 2] = nx[k].z; m[4 * k + 3] = nx[k].w; }
			if (done + 128 <= total) {
				const uint4 *p = reinterpret_cast<const uint4 *>(cu.ptr());
#pragma unroll
				for (int k = 0; k < 4; ++k) nx[k] = __ldg(p + k);
				cu.advance(64);
			}
			md5_block(st, m);
		}
	} else {
		for (; done + 64 <= total; done += 64) {
#pragma unroll
			for (int k = 0; k < 16; ++k) {
				uint32_t v = 0;
				for (int b = 0; b < 4; ++b) { v |= (uint32_t)__ldg(cu.ptr()) << (8 * b); cu.advance(1); }
				m[k] = v;
			}
			md5_block(st, m);
		}
	}
	// tail: the remaining bytes, 0x80, zeros, the length in bits
	const int rem = (int)(total - done);
#pragma unroll
	for (int k = 0; k < 16; ++k) m[k] = 0;
	for (int b = 0; b < rem; ++b) { m[b >> 2] |= (uint32_t)__ldg(cu.ptr()) << (8 * (b & 3)); cu.advance(1); }
	m[rem >> 2] |= 0x80u << (8 * (rem & 3));
	if (rem >= 56) {
		md5_block(st, m);
#pragma unroll
		for (int k = 0; k < 16; ++k) m[k] = 0;
	}
	m[14] = (uint32_t)(total << 3);
	m[15] = (uint32_t)((total << 3) >> 32);
	md5_block(st, m);
#pragma unroll
	for (int k = 0; k < 4; ++k) reinterpret_cast<uint32_t *>(out)[4 * i + k] = st[k];
}

}  // namespace

// digests[16 * i] = MD5 of the window (left, top, w, h) of frame slots[i] (index into `frames`, frame_bytes apart)
int md5_pictures(cudaStream_t stream, const uint8_t *frames, size_t frame_bytes, const int *d_slots, int n, int stride, int rows, int left, int top, int w, int h,
                 uint8_t *d_out) {
	if (n <= 0) return 0;
	md5_pictures_kernel<<<(n + 31) / 32, 32, 0, stream>>>(frames, frame_bytes, d_slots, 0, n, stride, (size_t)stride * rows, left, top, w, h, d_out);
	B2_CUDA_OK(cudaGetLastError());
	return 0;
}

// one frame (index `slot`), its planes addressed as the application does: row pitch `stride`, CbCr plane `plane_bytes`
// behind the luma plane
int md5_frame(cudaStream_t stream, const uint8_t *frames, size_t frame_bytes, int slot, int stride, size_t plane_bytes, int left, int top, int w, int h, uint8_t *d_out) {
	md5_pictures_kernel<<<1, 32, 0, stream>>>(frames, frame_bytes, nullptr, slot, 1, stride, plane_bytes, left, top, w, h, d_out);
	B2_CUDA_OK(cudaGetLastError());
	return 0;
}

}  // namespace b2
