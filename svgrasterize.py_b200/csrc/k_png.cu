// k_png.cu -- PNG encoding of the rendered RGBA8 canvases on the device.
//
// Replaces canvas_to_png (svgrasterize.py:249-274: filter-0 rows through zlib level 9 on one host core, 6.7 s
// for a 4096 x 4096 canvas) for the batch path, where the raw RGBA8 result (512 MiB per 2048 icons) is what
// binds the end-to-end step to PCIe: only the compressed files cross the bus.  The files are standard PNGs
// (8-bit RGBA, one IDAT) and decode to exactly the canvas bytes; they are not the reference's bytes (different
// filter and deflate strategy) -- the byte-identical path stays canvas_to_png's default.
//
//   filter     Paeth (type 4) on every row: flat areas and linear gradients become (near) zero residuals
//   LZ77       matches at distance 4 only (one pixel back in the residual stream), whole pixels, 1..64 pixels
//              per match: a run of equal residual pixels -- what flat, anti-aliased vector art consists of
//   Huffman    one dynamic block per segment (<= 256 work items of <= 256 pixels of a row): histogram ->
//              length-limited code (15 bits) built by the segment's CTA -> canonical codes; the code lengths go
//              out with a fixed 4-bit code-length code (no run-length symbols: 158 bytes of header per block)
//   framing    segments end on a byte boundary (an empty stored block, zlib's sync-flush marker), so they
//              concatenate by copying; zlib header, Adler-32 (combined from per-item sums), PNG chunks and their
//              CRC-32 (per-thread table CRC + GF(2) shifts) are written by the packing kernels
//
// One CTA per segment does all of its passes back to back (walk 1: histogram + Adler, Huffman code, walk 2: bit
// counts, scan, walk 3: emission); the filter is recomputed in every walk instead of being stored.
#include <algorithm>

#include "svgr_kernels.h"

#define PNG_THREADS 256
#define PNG_ITEM_PX 256         // pixels of a row per work item
#define PNG_NSYM 290            // 286 literal / length symbols + 4 distance symbols (codes 0..3)
#define PNG_MAXBITS 15
#define PNG_HEADER_BITS (3 + 5 + 5 + 4 + 19 * 3 + PNG_NSYM * 4)
#define ADLER_MOD 65521u

struct PngSeg {           // one deflate block
    int32_t canvas;       // index of the canvas
    int32_t row0, rows;   // rows of the canvas covered by the segment
    int32_t last;         // 1: last segment of its canvas (BFINAL)
    int64_t slot;         // byte offset of the segment's worst-case slot in the scratch buffer
};
struct PngCanvas {
    int64_t src;          // byte offset of the RGBA8 canvas in the canvas buffer
    int32_t rows, cols;
    int32_t seg0, nseg;   // its segments
};

namespace {

__constant__ unsigned char c_len_code[65];   // match of n pixels (4 n bytes): length symbol - 257
__constant__ unsigned char c_len_ebits[65];  // number of extra bits
__constant__ unsigned short c_len_extra[65]; // extra bits value
__constant__ unsigned c_crc_table[256];
__constant__ unsigned c_crc_pow[32];         // x^(8 * 2^k) mod P, reflected

__device__ __forceinline__ int paeth(int a, int b, int c)
{
    const int p = a + b - c;
    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// residual of pixel `cur` given its left, up and up-left neighbours (zero outside the image), bytes packed like
// the pixel
__device__ __forceinline__ unsigned residual(unsigned cur, unsigned left, unsigned up, unsigned ul)
{
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int sh = 8 * k;
        const int x = (cur >> sh) & 255, a = (left >> sh) & 255, b = (up >> sh) & 255, c = (ul >> sh) & 255;
        r |= (unsigned)((x - paeth(a, b, c)) & 255) << sh;
    }
    return r;
}

// Walks one work item = pixels [c0, c1) of row `row` of a canvas and feeds its tokens to `sink`:
//   sink.lit(byte), sink.match(n pixels).  ADLER: also accumulates the Adler-32 sums of the filtered bytes.
template <bool ADLER, class Sink>
__device__ __forceinline__ void walk_item(const unsigned *__restrict__ img, int cols, int row, int c0, int c1, Sink &sink,
                                          unsigned &ad_a, unsigned &ad_b)
{
    const unsigned *cur_row = img + (size_t)row * cols;
    const unsigned *up_row = row > 0 ? cur_row - cols : nullptr;
    unsigned left = c0 > 0 ? cur_row[c0 - 1] : 0u;
    unsigned ul = (up_row && c0 > 0) ? up_row[c0 - 1] : 0u;
    if (c0 == 0) {
        sink.lit(4);  // filter type byte: Paeth
        if (ADLER)
            ad_a += 4, ad_b += ad_a;
    }
    unsigned prev = 0;
    int run = 0;  // pixels of the pending match
    for (int c = c0; c < c1; c++) {
        const unsigned cur = cur_row[c], up = up_row ? up_row[c] : 0u;
        const unsigned r = residual(cur, left, up, ul);
        left = cur, ul = up;
        if (ADLER) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                ad_a += (r >> (8 * k)) & 255;
                ad_b += ad_a;
            }
        }
        if (c > c0 && r == prev) {
            if (++run == 64) {
                sink.match(64);
                run = 0;
            }
            continue;
        }
        if (run) {
            sink.match(run);
            run = 0;
        }
        sink.lit(r & 255), sink.lit((r >> 8) & 255), sink.lit((r >> 16) & 255), sink.lit(r >> 24);
        prev = r;
    }
    if (run)
        sink.match(run);
}

struct HistSink {
    unsigned *h;  // this warp's histogram
    __device__ __forceinline__ void lit(unsigned b) { atomicAdd(h + b, 1u); }
    __device__ __forceinline__ void match(int n)
    {
        atomicAdd(h + 257 + c_len_code[n], 1u);
        atomicAdd(h + 286 + 3, 1u);
    }
};
struct BitsSink {
    const unsigned char *len;
    unsigned bits;
    __device__ __forceinline__ void lit(unsigned b) { bits += len[b]; }
    __device__ __forceinline__ void match(int n) { bits += len[257 + c_len_code[n]] + c_len_ebits[n] + len[286 + 3]; }
};
struct EmitSink {
    const unsigned char *len;
    const unsigned short *code;
    unsigned *out;             // the segment's slot as 32-bit words (zeroed)
    unsigned long long acc;    // pending bits, LSB first
    int nacc;
    long long word;            // next word to write
    bool first;
    __device__ __forceinline__ void put(unsigned v, int n)
    {
        acc |= (unsigned long long)v << nacc;
        nacc += n;
        if (nacc >= 32) {
            const unsigned w = (unsigned)acc;
            if (first)
                atomicOr(out + word, w), first = false;  // shared with the item before
            else
                out[word] = w;
            word++, acc >>= 32, nacc -= 32;
        }
    }
    __device__ __forceinline__ void lit(unsigned b) { put(code[b], len[b]); }
    __device__ __forceinline__ void match(int n)
    {
        const int s = 257 + c_len_code[n];
        put(code[s], len[s]);
        if (c_len_ebits[n])
            put(c_len_extra[n], c_len_ebits[n]);
        put(code[286 + 3], len[286 + 3]);
    }
    __device__ __forceinline__ void finish()
    {
        if (nacc > 0)
            atomicOr(out + word, (unsigned)acc);  // shared with the item after
    }
};

__device__ __forceinline__ unsigned bit_reverse(unsigned v, int n) { return __brev(v) >> (32 - n); }

// (a, b) of block X followed by a block Y of `n` bytes (Adler-32 sums, each started from a = 1, b = 0)
__device__ __forceinline__ void adler_combine(unsigned &a, unsigned &b, unsigned ya, unsigned yb, unsigned long long n)
{
    const unsigned long long rem = n % ADLER_MOD;
    const unsigned long long nb = ((unsigned long long)b + yb + rem * ((a + ADLER_MOD - 1) % ADLER_MOD)) % ADLER_MOD;
    a = (a + ya + ADLER_MOD - 1) % ADLER_MOD;
    b = (unsigned)nb;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// one CTA per segment: histogram, code, bit counts, emission
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PNG_THREADS)
png_deflate_kernel(const PngSeg *__restrict__ segs, const PngCanvas *__restrict__ canvases,
                   const unsigned char *__restrict__ canvas_buf, unsigned char *__restrict__ scratch,
                   int *__restrict__ seg_bytes, unsigned *__restrict__ seg_adler /* a, b per segment */)
{
    __shared__ unsigned s_hist[PNG_THREADS / 32][PNG_NSYM + 6];
    __shared__ unsigned s_freq[PNG_NSYM];
    __shared__ unsigned char s_len[PNG_NSYM];
    __shared__ unsigned short s_code[PNG_NSYM];
    __shared__ unsigned short s_order[PNG_NSYM];     // symbols with freq > 0 by ascending (freq, symbol)
    __shared__ int s_parent[2 * PNG_NSYM];
    __shared__ unsigned s_weight[2 * PNG_NSYM];
    __shared__ unsigned s_item_bits[PNG_THREADS + 1];
    __shared__ unsigned s_item_a[PNG_THREADS], s_item_b[PNG_THREADS];
    __shared__ int s_nused[2];

    const int tid = threadIdx.x, warp = tid >> 5;
    const PngSeg seg = segs[blockIdx.x];
    const PngCanvas cv = canvases[seg.canvas];
    const unsigned *img = reinterpret_cast<const unsigned *>(canvas_buf + cv.src);
    const int items_per_row = (cv.cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX;
    const int n_items = seg.rows * items_per_row;  // <= PNG_THREADS by construction
    const bool live = tid < n_items;
    const int row = seg.row0 + (live ? tid / items_per_row : 0);
    const int c0 = live ? (tid % items_per_row) * PNG_ITEM_PX : 0;
    const int c1 = min(cv.cols, c0 + PNG_ITEM_PX);

    for (int i = tid; i < (PNG_THREADS / 32) * (PNG_NSYM + 6); i += PNG_THREADS)
        (&s_hist[0][0])[i] = 0;
    __syncthreads();
    // ---- walk 1: histogram (per warp) and Adler-32 sums (per item)
    unsigned ad_a = 1, ad_b = 0;
    if (live) {
        HistSink hs{s_hist[warp]};
        walk_item<true>(img, cv.cols, row, c0, c1, hs, ad_a, ad_b);
    }
    s_item_a[tid] = ad_a % ADLER_MOD, s_item_b[tid] = ad_b % ADLER_MOD;
    __syncthreads();
    for (int s = tid; s < PNG_NSYM; s += PNG_THREADS) {
        unsigned f = 0;
#pragma unroll
        for (int w = 0; w < PNG_THREADS / 32; w++)
            f += s_hist[w][s];
        if (s == 256)
            f = 1;  // end of block
        if (s == 286 + 3 && f == 0)
            f = 1;  // at least one distance code must be described
        if (s >= 286 && s != 286 + 3)
            f = 0;
        s_freq[s] = f;
        s_len[s] = 0;
    }
    __syncthreads();
    // ---- Huffman code lengths of the literal / length alphabet (0..285); the distance alphabet has one code
    // order the used symbols by (freq, symbol): rank = number of used symbols that sort before this one
    for (int s = tid; s < 286; s += PNG_THREADS) {
        const unsigned f = s_freq[s];
        if (f) {
            int rank = 0;
            for (int q = 0; q < 286; q++) {
                const unsigned g = s_freq[q];
                rank += (g && (g < f || (g == f && q < s))) ? 1 : 0;
            }
            s_order[rank] = (unsigned short)s;
        }
    }
    if (tid == 0) {
        int n = 0;
        for (int q = 0; q < 286; q++)
            n += s_freq[q] ? 1 : 0;
        s_nused[0] = n;
    }
    __syncthreads();
    if (tid == 0) {
        const int n = s_nused[0];
        if (n == 1) {
            s_len[s_order[0]] = 1;  // cannot happen (end of block + the filter byte are two symbols), kept for safety
        } else {
            // two-queue Huffman: leaves 0..n-1 in ascending weight, internal nodes n..2n-2 are created in
            // ascending weight too
            for (int i = 0; i < n; i++)
                s_weight[i] = s_freq[s_order[i]];
            int leaf = 0, node = n, next = n;
            auto take = [&]() {
                if (leaf < n && (node >= next || s_weight[leaf] <= s_weight[node]))
                    return leaf++;
                return node++;
            };
            for (; next < 2 * n - 1; next++) {
                const int a = take(), b = take();
                s_weight[next] = s_weight[a] + s_weight[b];
                s_parent[a] = next, s_parent[b] = next;
            }
            // depths, root = 2n - 2; s_weight is reused for them
            s_weight[2 * n - 2] = 0;
            int count[PNG_MAXBITS + 2];
            for (int i = 0; i <= PNG_MAXBITS + 1; i++)
                count[i] = 0;
            for (int i = 2 * n - 3; i >= 0; i--) {
                const unsigned d = s_weight[s_parent[i]] + 1;
                s_weight[i] = d;
                if (i < n)
                    count[d > PNG_MAXBITS ? PNG_MAXBITS : d]++;
            }
            // enforce the length limit on the counts per length, keeping the code complete
            unsigned total = 0;
            for (int i = PNG_MAXBITS; i > 0; i--)
                total += (unsigned)count[i] << (PNG_MAXBITS - i);
            while (total != (1u << PNG_MAXBITS)) {
                count[PNG_MAXBITS]--;
                for (int i = PNG_MAXBITS - 1; i > 0; i--)
                    if (count[i]) {
                        count[i]--;
                        count[i + 1] += 2;
                        break;
                    }
                total--;
            }
            // the least frequent symbols get the longest codes
            int idx = 0;
            for (int l = PNG_MAXBITS; l > 0; l--)
                for (int q = count[l]; q > 0; q--)
                    s_len[s_order[idx++]] = (unsigned char)l;
        }
        s_len[286 + 3] = 1;  // distance 4: the only distance code, one bit
        // canonical codes, stored bit-reversed (deflate packs Huffman codes starting from their most significant bit)
        int bl_count[PNG_MAXBITS + 1];
        for (int i = 0; i <= PNG_MAXBITS; i++)
            bl_count[i] = 0;
        for (int q = 0; q < 286; q++)
            bl_count[s_len[q]]++;
        bl_count[0] = 0;
        unsigned next_code[PNG_MAXBITS + 1];
        unsigned code = 0;
        for (int b = 1; b <= PNG_MAXBITS; b++) {
            code = (code + bl_count[b - 1]) << 1;
            next_code[b] = code;
        }
        for (int q = 0; q < 286; q++) {
            const int l = s_len[q];
            s_code[q] = l ? (unsigned short)bit_reverse(next_code[l]++, l) : 0;
        }
        for (int q = 286; q < PNG_NSYM; q++)
            s_code[q] = 0;  // distance code 3 = "0"
    }
    __syncthreads();
    // ---- walk 2: bits per item, then their offsets behind the block header
    unsigned bits = 0;
    if (live) {
        BitsSink bs{s_len, 0};
        unsigned da = 0, db = 0;
        walk_item<false>(img, cv.cols, row, c0, c1, bs, da, db);
        bits = bs.bits;
    }
    s_item_bits[tid] = bits;
    __syncthreads();
    if (tid == 0) {
        // exclusive scan (<= 256 entries) + Adler-32 of the segment, items in stream order
        unsigned off = PNG_HEADER_BITS;
        unsigned a = 1, b = 0;
        for (int i = 0; i < n_items; i++) {
            const unsigned t = s_item_bits[i];
            s_item_bits[i] = off;
            off += t;
            const int ic0 = (i % items_per_row) * PNG_ITEM_PX;
            const unsigned long long nbytes = 4ull * (min(cv.cols, ic0 + PNG_ITEM_PX) - ic0) + (ic0 == 0 ? 1 : 0);
            adler_combine(a, b, s_item_a[i], s_item_b[i], nbytes);
        }
        s_item_bits[PNG_THREADS] = off;  // end of the items: the end-of-block code follows
        seg_adler[2 * blockIdx.x] = a, seg_adler[2 * blockIdx.x + 1] = b;
    }
    __syncthreads();
    const unsigned end_bits = s_item_bits[PNG_THREADS];
    // block = header + items + end of block [+ 3 bits of an empty stored block, padding to a byte, 00 00 ff ff]
    const unsigned eob_len = s_len[256];
    unsigned total_bits = end_bits + eob_len;
    unsigned total_bytes;
    if (seg.last) {
        total_bytes = (total_bits + 7) / 8;
    } else {
        total_bits += 3;
        total_bytes = (total_bits + 7) / 8 + 4;
    }
    unsigned *out = reinterpret_cast<unsigned *>(scratch + seg.slot);
    const unsigned n_words = (total_bytes + 3) / 4 + 1;
    for (unsigned i = tid; i < n_words; i += PNG_THREADS)
        out[i] = 0;
    __syncthreads();
    // ---- walk 3: emission
    if (live) {
        const unsigned off = s_item_bits[tid];
        EmitSink es{s_len, s_code, out, 0ull, (int)(off & 31), (long long)(off >> 5), true};
        unsigned da = 0, db = 0;
        walk_item<false>(img, cv.cols, row, c0, c1, es, da, db);
        es.finish();
    }
    if (tid == PNG_THREADS - 1 || (tid == 0 && PNG_THREADS == 1)) {
        // header: BFINAL, BTYPE = 2, HLIT = 29 (286 codes), HDIST = 3 (4 codes), HCLEN = 15 (19 lengths), the
        // code-length code (4 bits for 0..15, unused 16..18), then the 290 code lengths as 4-bit codes
        EmitSink hs{s_len, s_code, out, 0ull, 0, 0ll, true};
        hs.put(seg.last ? 1u : 0u, 1), hs.put(2u, 2), hs.put(29u, 5), hs.put(3u, 5), hs.put(15u, 4);
        const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < 19; i++)
            hs.put(order[i] < 16 ? 4u : 0u, 3);
        for (int q = 0; q < PNG_NSYM; q++)
            hs.put(bit_reverse(s_len[q], 4), 4);
        hs.finish();
        // end of block (+ sync marker)
        EmitSink ts{s_len, s_code, out, 0ull, (int)(end_bits & 31), (long long)(end_bits >> 5), true};
        ts.put(s_code[256], (int)eob_len);
        if (!seg.last)
            ts.put(0u, 3);  // BFINAL = 0, BTYPE = 0: stored block of length 0
        ts.finish();
        if (!seg.last) {
            unsigned char *tail = scratch + seg.slot + total_bytes - 4;
            tail[0] = 0, tail[1] = 0, tail[2] = 0xff, tail[3] = 0xff;
        }
        seg_bytes[blockIdx.x] = (int)total_bytes;
    }
}

// ---------------------------------------------------------------------------------------------
// packing: file sizes -> offsets (scan on the host side of this file), then one CTA per canvas writes its file
//   signature, IHDR, IDAT (zlib header, segments, Adler-32), IEND -- with the chunk CRCs
// ---------------------------------------------------------------------------------------------
__global__ void png_sizes_kernel(const PngCanvas *__restrict__ canvases, int n_canvas, const int *__restrict__ seg_bytes,
                                 int *__restrict__ file_bytes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_canvas)
        return;
    long long n = 8 + 25 + 12 + 2 + 4 + 12;  // signature, IHDR chunk, IDAT framing, zlib header, Adler-32, IEND chunk
    for (int s = 0; s < canvases[i].nseg; s++)
        n += seg_bytes[canvases[i].seg0 + s];
    file_bytes[i] = (int)n;
}

namespace {

__device__ __forceinline__ unsigned crc_mul(unsigned a, unsigned b)
{
    // product of two polynomials modulo the CRC-32 polynomial, reflected bit order
    unsigned p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u)
            p ^= b;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}

// crc (zero-initialised, no final xor) of a message followed by n zero bytes
__device__ __forceinline__ unsigned crc_shift(unsigned crc, unsigned long long n)
{
    for (int k = 0; n; k++, n >>= 1)
        if (n & 1)
            crc = crc_mul(c_crc_pow[k], crc);
    return crc;
}

__device__ __forceinline__ void put_be32(unsigned char *p, unsigned v)
{
    p[0] = (unsigned char)(v >> 24), p[1] = (unsigned char)(v >> 16), p[2] = (unsigned char)(v >> 8), p[3] = (unsigned char)v;
}

// CRC-32 (zero init, no final xor) of bytes [0, n) by the whole CTA; result valid in thread 0
__device__ unsigned block_crc0(const unsigned char *p, long long n, unsigned *s_part, long long *s_len)
{
    const int tid = threadIdx.x;
    const long long span = ((n + PNG_THREADS - 1) / PNG_THREADS + 3) & ~3ll;
    const long long a = min(n, (long long)tid * span), b = min(n, a + span);
    unsigned crc = 0;
    for (long long i = a; i < b; i++)
        crc = c_crc_table[(crc ^ p[i]) & 255] ^ (crc >> 8);
    s_part[tid] = crc, s_len[tid] = b - a;
    __syncthreads();
    for (int step = 1; step < PNG_THREADS; step <<= 1) {
        if ((tid & (2 * step - 1)) == 0) {
            const long long lb = s_len[tid + step];
            if (lb > 0) {
                s_part[tid] = crc_shift(s_part[tid], (unsigned long long)lb) ^ s_part[tid + step];
                s_len[tid] += lb;
            }
        }
        __syncthreads();
    }
    return s_part[0];
}

}  // namespace

__global__ void __launch_bounds__(PNG_THREADS)
png_pack_kernel(const PngCanvas *__restrict__ canvases, const PngSeg *__restrict__ segs, const int *__restrict__ seg_bytes,
                const unsigned *__restrict__ seg_adler, const unsigned char *__restrict__ scratch,
                const long long *__restrict__ file_off, unsigned char *__restrict__ out)
{
    __shared__ unsigned s_part[PNG_THREADS];
    __shared__ long long s_len[PNG_THREADS];
    const int tid = threadIdx.x;
    const PngCanvas cv = canvases[blockIdx.x];
    unsigned char *f = out + file_off[blockIdx.x];
    long long idat = 2 + 4;
    for (int s = 0; s < cv.nseg; s++)
        idat += seg_bytes[cv.seg0 + s];
    unsigned char *body = f + 8 + 25 + 8;  // IDAT data
    if (tid == 0) {
        const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
        for (int i = 0; i < 8; i++)
            f[i] = sig[i];
        unsigned char *h = f + 8;
        put_be32(h, 13);
        h[4] = 'I', h[5] = 'H', h[6] = 'D', h[7] = 'R';
        put_be32(h + 8, (unsigned)cv.cols), put_be32(h + 12, (unsigned)cv.rows);
        h[16] = 8, h[17] = 6, h[18] = 0, h[19] = 0, h[20] = 0;
        unsigned crc = 0xffffffffu;
        for (int i = 4; i < 21; i++)
            crc = c_crc_table[(crc ^ h[i]) & 255] ^ (crc >> 8);
        put_be32(h + 21, ~crc);
        unsigned char *d = f + 8 + 25;
        put_be32(d, (unsigned)idat);
        d[4] = 'I', d[5] = 'D', d[6] = 'A', d[7] = 'T';
        body[0] = 0x78, body[1] = 0x5e;  // zlib: deflate, 32 KiB window, "fast" level hint
        // Adler-32 of the whole filtered image from the segments' sums
        unsigned a = 1, b = 0;
        const int items_per_row = (cv.cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX;
        (void)items_per_row;
        for (int s = 0; s < cv.nseg; s++) {
            const unsigned long long nbytes = (unsigned long long)segs[cv.seg0 + s].rows * (4ull * cv.cols + 1);
            adler_combine(a, b, seg_adler[2 * (cv.seg0 + s)], seg_adler[2 * (cv.seg0 + s) + 1], nbytes);
        }
        put_be32(body + idat - 4, (b << 16) | a);
        unsigned char *e = body + idat + 4;  // behind the IDAT crc
        put_be32(e, 0);
        e[4] = 'I', e[5] = 'E', e[6] = 'N', e[7] = 'D';
        put_be32(e + 8, 0xAE426082u);
    }
    // segments -> body (byte copies; the segments are byte aligned)
    long long pos = 2;
    for (int s = 0; s < cv.nseg; s++) {
        const unsigned char *src = scratch + segs[cv.seg0 + s].slot;
        const int n = seg_bytes[cv.seg0 + s];
        for (int i = tid; i < n; i += PNG_THREADS)
            body[pos + i] = src[i];
        pos += n;
    }
    __syncthreads();
    // CRC of the IDAT chunk: type + data
    const unsigned c0 = block_crc0(body - 4, idat + 4, s_part, s_len);
    if (tid == 0) {
        // a CRC started from 0xffffffff = the zero-started one xor the shifted initial value
        const unsigned crc = c0 ^ crc_shift(0xffffffffu, (unsigned long long)(idat + 4));
        put_be32(body + idat, ~crc);
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static void png_init_tables()
{
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev])
        return;
    // deflate length codes (RFC 1951, 3.2.5)
    const int base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const int ebits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    unsigned char code[65], eb[65];
    unsigned short extra[65];
    code[0] = eb[0] = 0, extra[0] = 0;
    for (int n = 1; n <= 64; n++) {
        const int len = 4 * n;
        int c = 28;
        while (c > 0 && (base[c] > len || c == 28))
            c--;  // code 28 is the literal length 258, never used here
        code[n] = (unsigned char)c, eb[n] = (unsigned char)ebits[c], extra[n] = (unsigned short)(len - base[c]);
    }
    cudaMemcpyToSymbol(c_len_code, code, sizeof code);
    cudaMemcpyToSymbol(c_len_ebits, eb, sizeof eb);
    cudaMemcpyToSymbol(c_len_extra, extra, sizeof extra);
    unsigned table[256];
    for (unsigned i = 0; i < 256; i++) {
        unsigned c = i;
        for (int k = 0; k < 8; k++)
            c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
        table[i] = c;
    }
    cudaMemcpyToSymbol(c_crc_table, table, sizeof table);
    // x^(8 * 2^k) mod P in reflected order: x^8 is the CRC of the byte 0x80 followed by ... computed by squaring
    auto mul = [](unsigned a, unsigned b) {
        unsigned p = 0;
        for (int i = 0; i < 32; i++) {
            if (a & 0x80000000u)
                p ^= b;
            a <<= 1;
            b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
        }
        return p;
    };
    unsigned pow[32];
    unsigned x8 = 0x80000000u;  // the polynomial "1" in reflected order
    for (int i = 0; i < 8; i++)
        x8 = (x8 >> 1) ^ ((x8 & 1u) ? 0xEDB88320u : 0u);  // times x, eight times
    pow[0] = x8;
    for (int k = 1; k < 32; k++)
        pow[k] = mul(pow[k - 1], pow[k - 1]);
    cudaMemcpyToSymbol(c_crc_pow, pow, sizeof pow);
    done[dev] = true;
}

// Upper bound of the bytes of a segment of `rows` x `cols`: a Huffman code is never longer in total than the
// fixed 9-bit code of its 286 symbols and a match spends at most three symbols on at least four bytes, so
// 10 bits per filtered byte cover the tokens; plus header, end of block and the sync marker.
long long svgr_png_slot_bytes(int rows, int cols)
{
    const long long raw = (long long)rows * (4ll * cols + 1);
    return ((raw * 10 + 7) / 8 + (PNG_HEADER_BITS + 7) / 8 + 32 + 15) & ~15ll;
}

int svgr_png_rows_per_segment(int cols)
{
    const int items_per_row = (cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX;
    return std::max(1, PNG_THREADS / items_per_row);
}

void svgr_launch_png_deflate(const void *segs, int n_seg, const void *canvases, const unsigned char *canvas_buf,
                             unsigned char *scratch, int *seg_bytes, unsigned *seg_adler, cudaStream_t s)
{
    png_init_tables();
    if (n_seg > 0)
        png_deflate_kernel<<<n_seg, PNG_THREADS, 0, s>>>((const PngSeg *)segs, (const PngCanvas *)canvases, canvas_buf, scratch,
                                                         seg_bytes, seg_adler);
}

void svgr_launch_png_sizes(const void *canvases, int n_canvas, const int *seg_bytes, int *file_bytes, cudaStream_t s)
{
    if (n_canvas > 0)
        png_sizes_kernel<<<(n_canvas + 255) / 256, 256, 0, s>>>((const PngCanvas *)canvases, n_canvas, seg_bytes, file_bytes);
}

void svgr_launch_png_pack(const void *canvases, int n_canvas, const void *segs, const int *seg_bytes, const unsigned *seg_adler,
                          const unsigned char *scratch, const long long *file_off, unsigned char *out, cudaStream_t s)
{
    png_init_tables();
    if (n_canvas > 0)
        png_pack_kernel<<<n_canvas, PNG_THREADS, 0, s>>>((const PngCanvas *)canvases, (const PngSeg *)segs, seg_bytes, seg_adler,
                                                         scratch, file_off, out);
}
