// k_png.cu -- PNG encoding of the rendered RGBA8 canvases on the device.
//
// Replaces canvas_to_png (svgrasterize.py:249-274: filter-0 rows through zlib level 9 on one host core, 6.7 s
// for a 4096 x 4096 canvas) for the batch path, where the raw RGBA8 result (512 MiB per 2048 icons) is what
// binds the end-to-end step to PCIe: only the compressed files cross the bus.  The files are standard PNGs
// (8-bit RGBA, one IDAT) and decode to exactly the canvas bytes; they are not the reference's bytes (different
// filter and deflate strategy) -- the byte-identical path stays canvas_to_png's default.
//
//   filter     Sub (type 1) on every row: one packed byte subtraction per pixel, no second row to read; flat
//              areas become zeros, linear gradients near-constant small residuals.  (Paeth was measured first:
//              files 3 % smaller, encoder 4 x slower -- 80 instructions per pixel wherever a gradient defeats
//              its shortcuts.)
//   LZ77       matches at distance 4 only (one pixel back in the residual stream), whole pixels, 1..64 pixels
//              per match: a run of equal residual pixels -- what flat, anti-aliased vector art consists of
//   Huffman    one dynamic block per segment (<= 256 work items of <= 256 pixels of a row): histogram of every
//              fourth item -> length-limited code (15 bits) built by the segment's CTA -> canonical codes; the
//              code lengths go out with a fixed 4-bit code-length code (158 bytes of header per block).  A call
//              with many segments (a batch of icons) shares ONE code: png_hist_kernel samples every 16th item of
//              every segment, png_code_kernel builds the code once, the segments only emit (files 2 % larger,
//              encoder 12 % faster)
//   framing    segments end on a byte boundary (an empty stored block, zlib's sync-flush marker), so they
//              concatenate by copying; zlib header, Adler-32 (per-item sums by dp4a, combined), PNG chunks and
//              their CRC-32 (per-thread table CRC + GF(2) shifts) are written by the packing kernel
//
// One CTA per segment; a warp takes an item at a time, 8 pixels per lane (two 16-byte loads): tokens are looked
// up once and kept in registers as (bits, length) per pixel, a warp scan places the lanes' bits in a shared-memory
// staging slot, the slot goes to the item's private area; when all items are done their offsets are known and
// they are merged (shifted) behind the block header.
#include <stdlib.h>

#include <algorithm>
#include <atomic>

#include "svgr_kernels.h"

#define PNG_THREADS 256
#define PNG_WARPS (PNG_THREADS / 32)
#define PNG_ITEM_PX 256         // pixels of a row per work item
#define PNG_NSYM 290            // 286 literal / length symbols + 4 distance symbols (codes 0..3)
#define PNG_MAXBITS 15
#define PNG_HEADER_BITS (3 + 5 + 5 + 4 + 19 * 3 + PNG_NSYM * 4)
#define PNG_ITEM_WORDS 484      // words of an item's private slot: (4 * 256 + 1) bytes at <= 15 bits each (an item
                                // the sample did not see may consist of the rarest symbols only)
#define ADLER_MOD 65521u
#define PNG_SHARED_SAMPLE 16     // shared code: every how many items of a segment enter the call's histogram
#define PNG_SHARED_MIN_SEGS 64   // a call with at least this many segments shares one code

struct PngSeg {           // one deflate block
    int32_t canvas;       // index of the canvas
    int32_t row0, rows;   // rows of the canvas covered by the segment
    int32_t last;         // 1: last segment of its canvas (BFINAL)
    int64_t slot;         // byte offset of the segment's scratch (item slots, then the merged stream)
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
__constant__ unsigned c_crc_slice[4][256];  // slicing-by-4: [k][i] = CRC of byte i followed by k zero bytes
__constant__ unsigned c_crc_pow[32];         // x^(8 * 2^k) mod P, reflected

// LSB-first bit writer into 32-bit words that start zeroed; the first and the last word of a writer may be shared
// with its neighbours (atomicOr), the words in between are its own
struct BitWriter {
    unsigned *out;
    unsigned long long acc;
    int nacc;
    long long word;
    bool first;
    __device__ __forceinline__ void put(unsigned v, int n)  // n <= 32
    {
        acc |= (unsigned long long)v << nacc;
        nacc += n;
        if (nacc >= 32) {
            const unsigned w = (unsigned)acc;
            if (first)
                atomicOr(out + word, w), first = false;
            else
                out[word] = w;
            word++, acc >>= 32, nacc -= 32;
        }
    }
    __device__ __forceinline__ void put64(unsigned long long v, int n)  // n <= 64
    {
        if (n > 32) {
            put((unsigned)v, 32);
            put((unsigned)(v >> 32), n - 32);
        } else if (n > 0) {
            put((unsigned)v, n);
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nacc > 0)
            atomicOr(out + word, (unsigned)acc);
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

// The 8 pixels of one lane of a work item (a warp covers the item's <= 256 pixels): Sub-filter residuals and the
// token structure.
struct LanePixels {
    unsigned r[8];      // residuals, bytes packed like the pixel
    unsigned lit;       // bit j: pixel j is a literal pixel (its residual differs from the pixel before)
    unsigned valid;     // bit j: pixel j exists
    int s_in;           // item-local index of the last literal pixel before this lane's pixels
    unsigned next_lit;  // the pixel after this lane's last one is a literal (or the item ends there)
};

__device__ __forceinline__ void lane_pixels(const unsigned *__restrict__ img, int cols, int row, int c0, int c1, int lane,
                                            LanePixels &px)
{
    const unsigned *cur_row = img + (size_t)row * cols;
    const int p0 = c0 + 8 * lane;
    const int nv = max(0, min(8, c1 - p0));
    unsigned cur[8];
    if (nv == 8 && ((reinterpret_cast<size_t>(cur_row + p0) & 15) == 0)) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(cur_row + p0)), b = __ldg(reinterpret_cast<const uint4 *>(cur_row + p0) + 1);
        cur[0] = a.x, cur[1] = a.y, cur[2] = a.z, cur[3] = a.w, cur[4] = b.x, cur[5] = b.y, cur[6] = b.z, cur[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++)
            cur[j] = j < nv ? __ldg(cur_row + p0 + j) : 0u;
    }
    unsigned left = __shfl_up_sync(0xffffffffu, cur[7], 1);
    if (lane == 0)
        left = c0 > 0 ? __ldg(cur_row + c0 - 1) : 0u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        px.r[j] = __vsub4(cur[j], left);  // per byte, modulo 256
        left = cur[j];
    }
    unsigned prev = __shfl_up_sync(0xffffffffu, px.r[7], 1);
    px.valid = nv >= 8 ? 0xffu : ((1u << nv) - 1u);
    unsigned lit = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const bool first = lane == 0 && j == 0;  // nothing to repeat at the start of an item
        if (first || px.r[j] != prev)
            lit |= 1u << j;
        prev = px.r[j];
    }
    px.lit = lit & px.valid;
    // last literal pixel before this lane: inclusive max-scan of every lane's last literal index, shifted by one lane
    int last = px.lit ? 8 * lane + (31 - __clz((int)px.lit)) : -1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, last, o);
        if (lane >= o)
            last = max(last, t);
    }
    px.s_in = __shfl_up_sync(0xffffffffu, last, 1);
    if (lane == 0)
        px.s_in = -1;
    const unsigned mine = (px.valid & 1u) ? (px.lit & 1u) : 1u;
    px.next_lit = __shfl_down_sync(0xffffffffu, mine, 1);
    if (lane == 31)
        px.next_lit = 1u;
}

// Token of pixel j of a lane: 0 nothing (inside a run), -1 a literal pixel, n > 0 a match of n pixels ending here
// (a run of repeated residual pixels ends, or reaches 64 pixels).  `s` carries the last literal pixel's index.
__device__ __forceinline__ int pixel_token(const LanePixels &px, int lane, int j, int &s)
{
    const int i = 8 * lane + j;
    if (px.lit >> j & 1u) {
        s = i;
        return -1;
    }
    const int d = i - s;  // >= 1: pixels of the run so far
    const bool last_of_lane = j == 7 || !(px.valid >> (j + 1) & 1u);
    const bool run_ends = last_of_lane ? (j == 7 ? px.next_lit != 0u : true) : ((px.lit >> (j + 1)) & 1u) != 0u;
    return ((d & 63) == 0 || run_ends) ? ((d - 1) & 63) + 1 : 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// one CTA per segment: sampled histogram -> code -> every item emitted into its own slot (warp per item, staged in
// shared memory) -> items merged bit-exactly behind the block header into the segment's stream
// ---------------------------------------------------------------------------------------------
// Everything the Huffman code of a block consists of, in shared memory: frequencies in, lengths / codes / match
// tokens out (build_code).
struct CodeTables {
    unsigned freq[PNG_NSYM];
    unsigned char len[PNG_NSYM];
    unsigned short code[PNG_NSYM];
    unsigned short order[PNG_NSYM];     // symbols by ascending (freq, symbol)
    int parent[2 * PNG_NSYM];
    unsigned weight[2 * PNG_NSYM];
    int count[PNG_MAXBITS + 2];         // symbols per code length
    int n_unseen, pseudo_pos;           // symbols outside the sample; where their pseudo-leaf sorts among the others
    unsigned next_code[PNG_MAXBITS + 1];
    unsigned match_bits[65];            // code + extra + distance of a match of n pixels ...
    unsigned char match_len[65];        // ... and its length in bits
};

// What a whole call shares when its segments are many and alike (png_code_kernel -> png_deflate_kernel<true>)
struct PngCode {
    unsigned char len[PNG_NSYM];
    unsigned short code[PNG_NSYM];
    unsigned match_bits[65];
    unsigned char match_len[65];
};

// Sampled histogram of one segment's tokens into s_hist (one row per warp), `sample` = every how many items
__device__ __forceinline__ void segment_histogram(unsigned (*s_hist)[PNG_NSYM + 6], const unsigned *img, const PngCanvas &cv,
                                                  const PngSeg &seg, int items_per_row, int n_items, int sample)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int it = warp * sample; it < n_items; it += PNG_WARPS * sample) {
        const int row = seg.row0 + it / items_per_row, c0 = (it % items_per_row) * PNG_ITEM_PX;
        LanePixels px;
        lane_pixels(img, cv.cols, row, c0, min(cv.cols, c0 + PNG_ITEM_PX), lane, px);
        unsigned *h = s_hist[warp];
        if (c0 == 0 && lane == 0)
            atomicAdd(h + 1, 1u);  // filter type byte: Sub
        int s = px.s_in;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (px.valid >> j & 1u) {
                const int t = pixel_token(px, lane, j, s);
                if (t < 0) {
                    const unsigned r = px.r[j];
                    atomicAdd(h + (r & 255), 1u), atomicAdd(h + ((r >> 8) & 255), 1u);
                    atomicAdd(h + ((r >> 16) & 255), 1u), atomicAdd(h + (r >> 24), 1u);
                } else if (t > 0) {
                    atomicAdd(h + 257 + c_len_code[t], 1u);
                }
            }
    }
}

// Length-limited canonical Huffman code from T.freq (every thread of the CTA calls this after a barrier that made
// the frequencies visible; returns with the tables complete and a barrier behind them).
__device__ void build_code(CodeTables &T)
{
    const int tid = threadIdx.x;
    unsigned *s_freq = T.freq;
    unsigned char *s_len = T.len;
    unsigned short *s_code = T.code, *s_order = T.order;
    int *s_parent = T.parent, *s_count = T.count;
    unsigned *s_weight = T.weight, *s_next_code = T.next_code, *s_match_bits = T.match_bits;
    unsigned char *s_match_len = T.match_len;
    // ---- Huffman code lengths of the literal / length alphabet (0..285); the distance alphabet has one code.
    // Symbols the sample never saw (frequency 1 after the smoothing -- two thirds of the alphabet for an icon) do not
    // enter the merge one by one: they hang as a complete binary subtree below ONE pseudo-leaf of their total weight,
    // which is where Huffman's algorithm would put most of them anyway; the sequential merge shrinks from 285 steps
    // to the number of symbols that occur.
    if (tid == 0)
        T.n_unseen = 0, T.pseudo_pos = 0;
    __syncthreads();
    for (int s = tid; s < 286; s += PNG_THREADS) {
        const unsigned f = s_freq[s];
        int rank = 0;
        for (int q = 0; q < 286; q++) {
            const unsigned g = s_freq[q];
            rank += (g < f || (g == f && q < s)) ? 1 : 0;
        }
        s_order[rank] = (unsigned short)s;
        if (f == 1)
            atomicAdd(&T.n_unseen, 1);
    }
    __syncthreads();
    const int k = T.n_unseen;                    // the first k symbols of the ascending order
    const int n_seen = 286 - k;
    const int m = n_seen + (k > 0 ? 1 : 0);      // leaves of the merge
    for (int s = tid; s < 286; s += PNG_THREADS) // the pseudo-leaf (weight k) sorts behind the seen symbols lighter than it
        if (s_freq[s] > 1 && s_freq[s] < (unsigned)k)
            atomicAdd(&T.pseudo_pos, 1);
    __syncthreads();
    const int ppos = k > 0 ? T.pseudo_pos : -1;
    for (int r = tid; r < n_seen; r += PNG_THREADS)
        s_weight[r + ((k > 0 && r >= ppos) ? 1 : 0)] = s_freq[s_order[k + r]];
    if (tid == 0 && k > 0)
        s_weight[ppos] = (unsigned)k;
    __syncthreads();
    if (tid == 0 && m >= 2) {
        const int n = m;
        // two-queue Huffman: leaves 0..n-1 in ascending weight; internal nodes n..2n-2 come out ascending too.
        // This merge is the one sequential piece of the construction; everything after it runs on all threads.
        int leaf = 0, node = n, next = n;
        unsigned w_leaf = s_weight[0], w_node = 0xffffffffu;  // heads of the two queues, kept in registers
        auto take = [&](unsigned &w) {
            if (leaf < n && (node >= next || w_leaf <= w_node)) {
                w = w_leaf;
                const int r = leaf++;
                w_leaf = leaf < n ? s_weight[leaf] : 0xffffffffu;
                return r;
            }
            w = w_node;
            const int r = node++;
            w_node = node < next ? s_weight[node] : 0xffffffffu;
            return r;
        };
        for (; next < 2 * n - 1; next++) {
            unsigned wa, wb;
            const int a = take(wa);
            const int b = take(wb);
            const unsigned w = wa + wb;
            s_weight[next] = w;
            if (node == next)
                w_node = w;  // the node queue was empty: the new node is its head
            s_parent[a] = next, s_parent[b] = next;
        }
    }
    for (int i = tid; i <= PNG_MAXBITS + 1; i += PNG_THREADS)
        s_count[i] = 0;
    __syncthreads();
    // depth of every leaf (its code length before the limit): walk to the root, all walks side by side; the
    // pseudo-leaf contributes the depths of its complete subtree of k leaves
    for (int i = tid; i < m; i += PNG_THREADS) {
        int d = 0;
        if (m >= 2)
            for (int q = i; q != 2 * m - 2; q = s_parent[q])
                d++;
        if (i == ppos && k > 1) {
            const int e = 31 - __clz(k);  // floor(log2 k)
            const int shallow = (2 << e) - k, deep = 2 * (k - (1 << e));
            atomicAdd(&s_count[min(d + e, PNG_MAXBITS)], shallow);
            if (deep)
                atomicAdd(&s_count[min(d + e + 1, PNG_MAXBITS)], deep);
        } else {
            atomicAdd(&s_count[min(d, PNG_MAXBITS)], 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        // enforce the length limit on the counts per length, keeping the code complete
        unsigned total = 0;
        for (int i = PNG_MAXBITS; i > 0; i--)
            total += (unsigned)s_count[i] << (PNG_MAXBITS - i);
        while (total != (1u << PNG_MAXBITS)) {
            s_count[PNG_MAXBITS]--;
            for (int i = PNG_MAXBITS - 1; i > 0; i--)
                if (s_count[i]) {
                    s_count[i]--;
                    s_count[i + 1] += 2;
                    break;
                }
            total--;
        }
        // first canonical code of every length
        unsigned code = 0;
        s_count[0] = 0;
        for (int b = 1; b <= PNG_MAXBITS; b++) {
            code = (code + (unsigned)s_count[b - 1]) << 1;
            s_next_code[b] = code;
        }
        s_len[286 + 3] = 1;  // distance 4: the only distance code, one bit
    }
    __syncthreads();
    // the least frequent symbols get the longest codes: position p of the ascending order has length l where the
    // counts of the lengths above l cover fewer than p + 1 symbols
    for (int p = tid; p < 286; p += PNG_THREADS) {
        int l = PNG_MAXBITS, covered = s_count[PNG_MAXBITS];
        while (covered <= p)
            covered += s_count[--l];
        s_len[s_order[p]] = (unsigned char)l;
    }
    __syncthreads();
    // canonical codes, stored bit-reversed (deflate packs Huffman codes from their most significant bit): the
    // code of a symbol = first code of its length + the number of smaller symbols of that length
    for (int q = tid; q < PNG_NSYM; q += PNG_THREADS) {
        unsigned short c = 0;
        if (q < 286) {
            const int l = s_len[q];
            int before = 0;
            for (int r = 0; r < q; r++)
                before += s_len[r] == l ? 1 : 0;
            c = (unsigned short)bit_reverse(s_next_code[l] + (unsigned)before, l);
        }
        s_code[q] = c;  // distance code 3 = "0"
    }
    __syncthreads();
    if (tid >= 1 && tid <= 64) {
        // a match of tid pixels: length code, its extra bits, the one-bit distance code "0"
        const int sym = 257 + c_len_code[tid];
        const int l = s_len[sym], eb = c_len_ebits[tid];
        s_match_bits[tid] = (unsigned)s_code[sym] | ((unsigned)c_len_extra[tid] << l);
        s_match_len[tid] = (unsigned char)(l + eb + 1);
    }
    __syncthreads();
}

// Per call, when its segments are many and alike (a batch of icons): one sampled histogram over all segments ...
__global__ void __launch_bounds__(PNG_THREADS)
png_hist_kernel(const PngSeg *__restrict__ segs, const PngCanvas *__restrict__ canvases,
                const unsigned char *__restrict__ canvas_buf, unsigned *__restrict__ hist /* PNG_NSYM, zeroed */)
{
    __shared__ unsigned s_hist[PNG_WARPS][PNG_NSYM + 6];
    const int tid = threadIdx.x;
    const PngSeg seg = segs[blockIdx.x];
    const PngCanvas cv = canvases[seg.canvas];
    const unsigned *img = reinterpret_cast<const unsigned *>(canvas_buf + cv.src);
    const int items_per_row = (cv.cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX;
    for (int i = tid; i < PNG_WARPS * (PNG_NSYM + 6); i += PNG_THREADS)
        (&s_hist[0][0])[i] = 0;
    __syncthreads();
    segment_histogram(s_hist, img, cv, seg, items_per_row, seg.rows * items_per_row, PNG_SHARED_SAMPLE);
    __syncthreads();
    for (int s = tid; s < PNG_NSYM; s += PNG_THREADS) {
        unsigned f = 0;
#pragma unroll
        for (int w = 0; w < PNG_WARPS; w++)
            f += s_hist[w][s];
        if (f)
            atomicAdd(hist + s, f);
    }
}

// ... and one code for all of them (a single CTA)
__global__ void __launch_bounds__(PNG_THREADS)
png_code_kernel(const unsigned *__restrict__ hist, PngCode *__restrict__ out)
{
    __shared__ CodeTables T;
    const int tid = threadIdx.x;
    for (int s = tid; s < PNG_NSYM; s += PNG_THREADS) {
        T.freq[s] = s < 286 ? hist[s] + 1 : 0;  // every literal / length symbol can occur in the unsampled items
        T.len[s] = 0;
    }
    __syncthreads();
    build_code(T);
    for (int s = tid; s < PNG_NSYM; s += PNG_THREADS)
        out->len[s] = T.len[s], out->code[s] = T.code[s];
    if (tid <= 64)
        out->match_bits[tid] = tid ? T.match_bits[tid] : 0u, out->match_len[tid] = tid ? T.match_len[tid] : 0;
}

// SHARED_CODE: the Huffman code comes from png_code_kernel (the call's segments share it) instead of being built
// from this segment's own sampled histogram.
template <bool SHARED_CODE>
__global__ void __launch_bounds__(PNG_THREADS)
png_deflate_kernel(const PngSeg *__restrict__ segs, const PngCanvas *__restrict__ canvases,
                   const unsigned char *__restrict__ canvas_buf, unsigned char *__restrict__ scratch,
                   int *__restrict__ seg_bytes, unsigned *__restrict__ seg_adler /* a, b per segment */,
                   const PngCode *__restrict__ shared_code)
{
    __shared__ unsigned s_hist[SHARED_CODE ? 1 : PNG_WARPS][PNG_NSYM + 6];
    __shared__ CodeTables T;
    __shared__ unsigned s_item_bits[PNG_THREADS + 1];
    __shared__ unsigned s_item_a[PNG_THREADS], s_item_b[PNG_THREADS];
    __shared__ unsigned s_stage[PNG_WARPS][PNG_ITEM_WORDS];
    __shared__ unsigned s_warp_tot[3][PNG_WARPS];
    unsigned char *s_len = T.len;
    unsigned short *s_code = T.code;
    unsigned *s_match_bits = T.match_bits;
    unsigned char *s_match_len = T.match_len;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PngSeg seg = segs[blockIdx.x];
    const PngCanvas cv = canvases[seg.canvas];
    const unsigned *img = reinterpret_cast<const unsigned *>(canvas_buf + cv.src);
    const int items_per_row = (cv.cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX;
    const int n_items = seg.rows * items_per_row;  // <= PNG_THREADS by construction
    unsigned *item_slots = reinterpret_cast<unsigned *>(scratch + seg.slot);
    unsigned *out = item_slots + (size_t)n_items * PNG_ITEM_WORDS;

    if (SHARED_CODE) {
        for (int s = tid; s < PNG_NSYM; s += PNG_THREADS)
            T.len[s] = shared_code->len[s], T.code[s] = shared_code->code[s];
        if (tid <= 64)
            T.match_bits[tid] = shared_code->match_bits[tid], T.match_len[tid] = shared_code->match_len[tid];
        __syncthreads();
    } else {
        for (int i = tid; i < PNG_WARPS * (PNG_NSYM + 6); i += PNG_THREADS)
            (&s_hist[0][0])[i] = 0;
        __syncthreads();
        // ---- histogram of a sample of the items (every fourth; all of a small segment): the code only has to be
        // good, not optimal, and every symbol gets a code anyway (+1 below)
        segment_histogram(s_hist, img, cv, seg, items_per_row, n_items, n_items >= 32 ? 4 : 1);
        __syncthreads();
        for (int s = tid; s < PNG_NSYM; s += PNG_THREADS) {
            unsigned f = 0;
#pragma unroll
            for (int w = 0; w < (SHARED_CODE ? 1 : PNG_WARPS); w++)
                f += s_hist[w][s];
            T.freq[s] = s < 286 ? f + 1 : 0;  // every literal / length symbol can occur in the unsampled items
            T.len[s] = 0;
        }
        __syncthreads();
        build_code(T);
    }
    // ---- every item: tokens -> (bits, length) per pixel in registers -> staged in shared memory -> its slot
    for (int it = warp; it < n_items; it += PNG_WARPS) {
        const int row = seg.row0 + it / items_per_row, c0 = (it % items_per_row) * PNG_ITEM_PX;
        const int c1 = min(cv.cols, c0 + PNG_ITEM_PX);
        LanePixels px;
        lane_pixels(img, cv.cols, row, c0, c1, lane, px);
        const int head = c0 == 0 ? 1 : 0;
        const unsigned n_bytes = 4u * (unsigned)(c1 - c0) + head;
        unsigned long long tb[8];
        int tl[8];
        unsigned lane_bits = 0, s1 = 0, s2 = 0;
        if (head && lane == 0) {
            lane_bits = s_len[1];
            s1 = 1, s2 = n_bytes;  // the filter byte (value 1) at index 0
        }
        int s = px.s_in;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            tb[j] = 0ull, tl[j] = 0;
            if (px.valid >> j & 1u) {
                const unsigned r = px.r[j];
                const int t = pixel_token(px, lane, j, s);
                if (t < 0) {
                    const unsigned b0 = r & 255, b1 = (r >> 8) & 255, b2 = (r >> 16) & 255, b3 = r >> 24;
                    const int l0 = s_len[b0], l1 = s_len[b1], l2 = s_len[b2], l3 = s_len[b3];
                    tb[j] = (unsigned long long)s_code[b0] | ((unsigned long long)s_code[b1] << l0) |
                            ((unsigned long long)s_code[b2] << (l0 + l1)) | ((unsigned long long)s_code[b3] << (l0 + l1 + l2));
                    tl[j] = l0 + l1 + l2 + l3;
                } else if (t > 0) {
                    tb[j] = s_match_bits[t], tl[j] = s_match_len[t];
                }
                lane_bits += tl[j];
                // Adler-32 sums: A = 1 + sum x, B = N + sum (N - index) x
                const unsigned sum = __dp4a(r, 0x01010101u, 0u), wsum = __dp4a(r, 0x03020100u, 0u);
                s1 += sum;
                s2 += (n_bytes - (head + 4u * (8 * lane + j))) * sum - wsum;
            }
        }
        unsigned incl = lane_bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        const unsigned item_bits = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned off = incl - lane_bits;
        const unsigned n_words = (item_bits + 31) >> 5;
        unsigned *stage = s_stage[warp];
        for (unsigned k = lane; k < n_words; k += 32)
            stage[k] = 0;
        __syncwarp();
        BitWriter bw{stage, 0ull, (int)(off & 31), (long long)(off >> 5), true};
        if (head && lane == 0)
            bw.put(s_code[1], s_len[1]);
#pragma unroll
        for (int j = 0; j < 8; j++)
            bw.put64(tb[j], tl[j]);
        bw.finish();
        __syncwarp();
        unsigned *slot = item_slots + (size_t)it * PNG_ITEM_WORDS;
        for (unsigned k = lane; k < n_words; k += 32)
            slot[k] = stage[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            s_item_bits[it] = item_bits;
            s_item_a[it] = (1u + s1) % ADLER_MOD, s_item_b[it] = (n_bytes + s2) % ADLER_MOD;
        }
        __syncwarp();
    }
    __syncthreads();
    {
        // exclusive scan of the item sizes (<= 256 = one per thread) + Adler-32 of the segment, items in stream order:
        // A = 1 + sum (a_i - 1), B = sum b_i + sum n_i * (A before item i - 1)   (zlib's adler32_combine, unrolled)
        const bool on = tid < n_items;
        const unsigned bits = on ? s_item_bits[tid] : 0u;
        const unsigned a1 = on ? (s_item_a[tid] + ADLER_MOD - 1) % ADLER_MOD : 0u;
        unsigned inc_bits = bits, inc_a = a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t0 = __shfl_up_sync(0xffffffffu, inc_bits, o), t1 = __shfl_up_sync(0xffffffffu, inc_a, o);
            if (lane >= o)
                inc_bits += t0, inc_a += t1;
        }
        if (lane == 31)
            s_warp_tot[0][warp] = inc_bits, s_warp_tot[1][warp] = inc_a;
        __syncthreads();
        unsigned base_bits = PNG_HEADER_BITS, base_a = 0;
        for (int w = 0; w < warp; w++)
            base_bits += s_warp_tot[0][w], base_a += s_warp_tot[1][w];
        const unsigned ex_bits = base_bits + inc_bits - bits;
        const unsigned ex_a = base_a + inc_a - a1;  // < 256 * 65521: no reduction needed before the product
        unsigned term = 0;
        if (on) {
            const int ic0 = (tid % items_per_row) * PNG_ITEM_PX;
            const unsigned long long nbytes = 4ull * (min(cv.cols, ic0 + PNG_ITEM_PX) - ic0) + (ic0 == 0 ? 1 : 0);
            term = (unsigned)((s_item_b[tid] + nbytes * (ex_a % ADLER_MOD)) % ADLER_MOD);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            term += __shfl_xor_sync(0xffffffffu, term, o);  // <= 32 * 65520
        if (lane == 0)
            s_warp_tot[2][warp] = term;
        __syncthreads();  // every thread has read its s_item_bits entry
        if (on)
            s_item_bits[tid] = ex_bits;
        if (tid == PNG_THREADS - 1) {
            s_item_bits[n_items] = base_bits + inc_bits;  // end of the items: the end-of-block code follows
            unsigned bsum = 0;
            for (int w = 0; w < PNG_WARPS; w++)
                bsum += s_warp_tot[2][w];
            seg_adler[2 * blockIdx.x] = (1u + (base_a + inc_a) % ADLER_MOD) % ADLER_MOD;
            seg_adler[2 * blockIdx.x + 1] = bsum % ADLER_MOD;
        }
    }
    __syncthreads();
    const unsigned end_bits = s_item_bits[n_items];
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
    const unsigned n_out_words = (total_bytes + 3) / 4 + 1;
    for (unsigned i = tid; i < n_out_words; i += PNG_THREADS)
        out[i] = 0;
    __syncthreads();
    // ---- merge: every item's bits move to their place in the stream (a shift by the offset's low 5 bits)
    for (int it = warp; it < n_items; it += PNG_WARPS) {
        const unsigned d0 = s_item_bits[it], nbits = s_item_bits[it + 1] - d0;
        const unsigned sh = d0 & 31, w0 = d0 >> 5, n_src = (nbits + 31) >> 5;
        const unsigned n_dst = (sh + nbits + 31) >> 5;  // destination words touched
        const unsigned *slot = item_slots + (size_t)it * PNG_ITEM_WORDS;
        for (unsigned k = lane; k < n_dst; k += 32) {
            const unsigned lo = k < n_src ? slot[k] : 0u;
            const unsigned hi = (k > 0 && sh) ? slot[k - 1] : 0u;
            const unsigned v = (lo << sh) | (sh ? hi >> (32 - sh) : 0u);
            if (k == 0 || k + 1 >= n_dst)
                atomicOr(out + w0 + k, v);  // shared with the neighbouring items
            else
                out[w0 + k] = v;
        }
    }
    if (tid == PNG_THREADS - 1) {
        // header: BFINAL, BTYPE = 2, HLIT = 29 (286 codes), HDIST = 3 (4 codes), HCLEN = 15 (19 lengths), the
        // code-length code (4 bits for 0..15, unused 16..18), then the 290 code lengths as 4-bit codes
        BitWriter hs{out, 0ull, 0, 0ll, true};
        hs.put(seg.last ? 1u : 0u, 1), hs.put(2u, 2), hs.put(29u, 5), hs.put(3u, 5), hs.put(15u, 4);
        const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < 19; i++)
            hs.put(order[i] < 16 ? 4u : 0u, 3);
        hs.finish();  // the 290 code lengths follow, written by all threads below
        // end of block (+ sync marker)
        BitWriter ts{out, 0ull, (int)(end_bits & 31), (long long)(end_bits >> 5), true};
        ts.put(s_code[256], (int)eob_len);
        if (!seg.last)
            ts.put(0u, 3);  // BFINAL = 0, BTYPE = 0: stored block of length 0
        ts.finish();
        seg_bytes[blockIdx.x] = (int)total_bytes;
    }
    for (int q = tid; q < PNG_NSYM; q += PNG_THREADS) {
        // code length q as a 4-bit code of the fixed code-length code, at its place in the header
        const unsigned pos = PNG_HEADER_BITS - PNG_NSYM * 4 + 4u * q;
        const unsigned long long v = (unsigned long long)bit_reverse(s_len[q], 4) << (pos & 31);
        atomicOr(out + (pos >> 5), (unsigned)v);
        if ((unsigned)(v >> 32))
            atomicOr(out + (pos >> 5) + 1, (unsigned)(v >> 32));
    }
    if (!seg.last) {
        __syncthreads();  // the marker bytes share a word with the last bits
        if (tid == 0) {
            unsigned char *tail = reinterpret_cast<unsigned char *>(out) + total_bytes - 4;
            tail[0] = 0, tail[1] = 0, tail[2] = 0xff, tail[3] = 0xff;
        }
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

// CRC-32 (zero init, no final xor) of bytes [0, n) by the whole CTA; result valid in thread 0.  Every thread takes
// a contiguous span (a multiple of 4 bytes), reads it as aligned words (the message may start at any byte: two
// words and a funnel shift; up to 3 bytes behind the message are read, the caller guarantees they exist), four
// bytes per step through four tables, then moves its CRC to the end of the message (crc * x^(8 * bytes after the
// span), a few GF(2) multiplications, all threads at once) -- the message CRC is the xor of those.
__device__ unsigned block_crc0(const unsigned char *p, long long n, unsigned *s_part, const unsigned (*s_slice)[256])
{
    const int tid = threadIdx.x;
    const long long span = ((n + PNG_THREADS - 1) / PNG_THREADS + 3) & ~3ll;
    const long long a = min(n, (long long)tid * span), b = min(n, a + span);
    unsigned crc = 0;
    const int mis = (int)(reinterpret_cast<size_t>(p + a) & 3);
    const unsigned *w = reinterpret_cast<const unsigned *>(p + a - mis);
    const long long full = (b - a) >> 2;
    unsigned lo = full > 0 || b > a ? w[0] : 0u;
    for (long long i = 0; i < full; i++) {
        const unsigned hi = mis ? w[i + 1] : 0u;
        const unsigned v = mis ? __funnelshift_r(lo, hi, 8 * mis) : lo;
        lo = mis ? hi : (i + 1 < full || (b - a) & 3 ? w[i + 1] : 0u);
        crc ^= v;
        crc = s_slice[3][crc & 255] ^ s_slice[2][(crc >> 8) & 255] ^ s_slice[1][(crc >> 16) & 255] ^ s_slice[0][crc >> 24];
    }
    for (long long i = a + 4 * full; i < b; i++)
        crc = s_slice[0][(crc ^ p[i]) & 255] ^ (crc >> 8);
    crc = crc_shift(crc, (unsigned long long)(n - b));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
    if ((tid & 31) == 0)
        s_part[tid >> 5] = crc;
    __syncthreads();
    unsigned r = 0;
    for (int k = 0; k < PNG_WARPS; k++)
        r ^= s_part[k];
    return r;
}

}  // namespace

__global__ void __launch_bounds__(PNG_THREADS)
png_pack_kernel(const PngCanvas *__restrict__ canvases, const PngSeg *__restrict__ segs, const int *__restrict__ seg_bytes,
                const unsigned *__restrict__ seg_adler, const unsigned char *__restrict__ scratch,
                const long long *__restrict__ file_off, unsigned char *__restrict__ out)
{
    __shared__ unsigned s_part[PNG_WARPS];
    __shared__ unsigned s_slice[4][256];  // data-dependent indices: shared memory, not the constant cache
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * 256; i += PNG_THREADS)
        (&s_slice[0][0])[i] = (&c_crc_slice[0][0])[i];
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
        const PngSeg sg = segs[cv.seg0 + s];
        const int seg_items = sg.rows * ((cv.cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX);
        const unsigned char *src = scratch + sg.slot + (size_t)seg_items * PNG_ITEM_WORDS * 4;
        const int n = seg_bytes[cv.seg0 + s];
        // the source is word aligned, the destination is not: every thread builds aligned destination words from two
        // source words (funnel shift); the ragged ends go byte by byte
        unsigned char *dst = body + pos;
        const int lead = (int)((4 - (reinterpret_cast<size_t>(dst) & 3)) & 3);  // bytes up to the first aligned word
        const int n_words = n > lead ? (n - lead) / 4 : 0;
        for (int i = tid; i < min(lead, n); i += PNG_THREADS)
            dst[i] = src[i];
        const unsigned *sw = reinterpret_cast<const unsigned *>(src);
        unsigned *dw = reinterpret_cast<unsigned *>(dst + lead);
        const int sh = 8 * (lead & 3);
        for (int i = tid; i < n_words; i += PNG_THREADS) {
            const int sb = lead + 4 * i;  // source byte of this destination word
            const unsigned lo = sw[sb >> 2], hi = sw[(sb >> 2) + 1];  // the slot has a spare word behind the stream
            dw[i] = sh ? __funnelshift_r(lo, hi, sh) : lo;
        }
        for (int i = lead + 4 * n_words + tid; i < n; i += PNG_THREADS)
            dst[i] = src[i];
        pos += n;
    }
    __syncthreads();
    // CRC of the IDAT chunk: type + data
    const unsigned c0 = block_crc0(body - 4, idat + 4, s_part, s_slice);  // the 4 CRC bytes + IEND follow: readable
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
    // per device; two threads racing here both upload the same constants, which is harmless
    static std::atomic<bool> done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev].load(std::memory_order_acquire))
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
    unsigned slice[4][256];
    for (unsigned i = 0; i < 256; i++) {
        slice[0][i] = table[i];
        for (int k = 1; k < 4; k++)
            slice[k][i] = (slice[k - 1][i] >> 8) ^ table[slice[k - 1][i] & 255];
    }
    cudaMemcpyToSymbol(c_crc_slice, slice, sizeof slice);
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
    done[dev].store(true, std::memory_order_release);
}

// Upper bound of the bytes of a segment of `rows` x `cols`: the code is built from a sample of the items, so the
// only bound that always holds is the code length limit, 15 bits per filtered byte (a match spends at most 36 bits
// on at least four bytes); plus header, end of block and the sync marker.  Scratch, not output: files are packed.
long long svgr_png_slot_bytes(int rows, int cols)
{
    const long long raw = (long long)rows * (4ll * cols + 1);
    const long long items = (long long)rows * ((cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX);
    // the items' private slots, then the merged stream
    return (items * PNG_ITEM_WORDS * 4 + (raw * 15 + 7) / 8 + (PNG_HEADER_BITS + 7) / 8 + 32 + 15) & ~15ll;
}

int svgr_png_rows_per_segment(int cols)
{
    const int items_per_row = (cols + PNG_ITEM_PX - 1) / PNG_ITEM_PX;
    return std::max(1, PNG_THREADS / items_per_row);
}

size_t svgr_png_code_scratch_bytes() { return ((PNG_NSYM * sizeof(unsigned) + 15) & ~(size_t)15) + sizeof(PngCode); }

// code_scratch: svgr_png_code_scratch_bytes() of device memory (the call's histogram and its code).  Returns the
// number of kernels launched.
int svgr_launch_png_deflate(const void *segs, int n_seg, const void *canvases, const unsigned char *canvas_buf,
                            unsigned char *scratch, int *seg_bytes, unsigned *seg_adler, void *code_scratch, cudaStream_t s)
{
    png_init_tables();
    if (n_seg <= 0)
        return 0;
    const PngSeg *sg = (const PngSeg *)segs;
    const PngCanvas *cv = (const PngCanvas *)canvases;
    static const bool no_shared = getenv("SVGR_PNG_PER_SEGMENT_CODE") != nullptr;
    if (n_seg >= PNG_SHARED_MIN_SEGS && code_scratch && !no_shared) {
        // many segments (a batch of icons): one sampled histogram and one Huffman code for the whole call instead of
        // a histogram pass and a code construction per segment
        unsigned *hist = (unsigned *)code_scratch;
        PngCode *code = (PngCode *)((char *)code_scratch + ((PNG_NSYM * sizeof(unsigned) + 15) & ~(size_t)15));
        cudaMemsetAsync(hist, 0, PNG_NSYM * sizeof(unsigned), s);
        png_hist_kernel<<<n_seg, PNG_THREADS, 0, s>>>(sg, cv, canvas_buf, hist);
        png_code_kernel<<<1, PNG_THREADS, 0, s>>>(hist, code);
        png_deflate_kernel<true><<<n_seg, PNG_THREADS, 0, s>>>(sg, cv, canvas_buf, scratch, seg_bytes, seg_adler, code);
        return 3;
    }
    png_deflate_kernel<false><<<n_seg, PNG_THREADS, 0, s>>>(sg, cv, canvas_buf, scratch, seg_bytes, seg_adler, nullptr);
    return 1;
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
