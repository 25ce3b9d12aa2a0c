/* blow5_kernels.cuh — sm_100a device code for SURVEY.md §8(f) N4: BLOW5 records decoded on the device, so that what
 * crosses PCIe is the file's own (compressed) bytes and the raw signal never exists on the host.
 *
 * What is replaced (paths relative to the f5c tree): the reader side of slow5lib as f5c uses it in read_slow5_single
 * (src/f5cio.c:421-470): slow5_get -> record decompression (slow5lib/src/slow5_press.c:921-1010, zlib inflate) ->
 * record parsing (slow5lib/src/slow5.c:2840-2930: read_id, read_group, digitisation, offset, range, sampling_rate,
 * len_raw_signal, raw_signal) -> signal decompression (svb-zd: slow5_press.c:1118-1170, streamvbyte_decode.c,
 * streamvbyte_zigzag.c:34-40) -> the int16 -> float widening of src/f5cio.c:461.
 *
 *   abea_inflate_kernel        one WARP per record, lane 0 decoding: a complete DEFLATE decoder (RFC 1951: stored,
 *                              fixed and dynamic Huffman blocks) behind the zlib wrapper (RFC 1950). Records are
 *                              independent streams, a stream is bit-serial, so the parallelism is across records — and
 *                              a record gets a warp of its own because 32 decoders in one warp are 32 different control
 *                              flows (first version, one thread per record: 565 ms per 1792 records, every branch
 *                              divergent). Huffman tables live in shared memory, one set per warp: a 10-bit direct
 *                              table for literal/length codes, an 8-bit one for distance codes, canonical count/symbol
 *                              arrays for the (rare) longer codes.
 *   abea_blow5_parse_kernel    one thread per record: the fixed fields in front of the signal (unaligned loads).
 *   abea_blow5_signal_kernel   one WARP per record: int16 samples, or svb-zd (2-bit keys -> byte counts -> warp scan ->
 *                              values -> zigzag -> warp scan of the deltas), or ex-zd (slow5_press.c:1262-1842: one byte
 *                              per zigzag delta, the deltas above 255 kept aside as "exceptions" whose positions and
 *                              values are two streamvbyte streams; every lane finds by binary search how many
 *                              exceptions lie before its sample), widened to the float samples the event detection
 *                              kernels read.
 *
 * Integer / byte work throughout: results are bit-identical to slow5lib's by construction (tests/test_blow5.py checks
 * against Python's zlib and a restatement of svb-zd, on files written by slow5lib itself).
 * Compiles unchanged for the CPU SIMT emulator (tests/simt).
 */
#pragma once

#define B5_REC_NONE 0
#define B5_REC_ZLIB 1
#define B5_SIG_NONE 0
#define B5_SIG_SVB_ZD 1
#define B5_SIG_EX_ZD 2

#define B5_OK 0
#define B5_ERR_DATA 1     /* malformed stream / record */
#define B5_ERR_OVERFLOW 2 /* the output capacity was too small (the host retries with more) */

#define B5_INFLATE_WARPS 8   /* records per CTA */
#define B5_LIT_BITS 10
#define B5_DIST_BITS 8

/* one record's descriptor: where its stored bytes are and where its decompressed image goes */
struct abea_b5rec_t {
    int64_t in_off;   /* first stored byte in d_b5in */
    int64_t out_off;  /* first byte of the decompressed record in d_b5out (== in_off when the records are not compressed) */
    int32_t in_len;
    int32_t out_cap;
};

/* what the parser finds in a decompressed record */
struct abea_b5hdr_t {
    double digitisation, offset, range, sampling_rate;
    int64_t sig_off;    /* first byte of the signal field, relative to the record */
    int64_t sig_bytes;  /* its size in bytes as stored */
    int32_t n_samples;  /* samples after signal decompression */
    int32_t status;
};

/* per-thread Huffman tables of the inflater (shared memory) */
struct b5_tables_t {
    uint16_t lit[1 << B5_LIT_BITS];   /* (symbol << 4) | code length, 0 = longer than B5_LIT_BITS bits */
    uint16_t dist[1 << B5_DIST_BITS];
    uint16_t lsym[288], dsym[32];     /* symbols in canonical order */
    uint16_t lcnt[16], dcnt[16];      /* codes per length */
    uint8_t len[320];                 /* code lengths being read (literal/length then distance) */
};

struct b5_bits_t {
    const uint8_t* in;
    int64_t n, pos;
    uint64_t buf;
    int cnt;
    bool bad;
};

__device__ __forceinline__ void b5_refill(b5_bits_t& b) {
    while (b.cnt <= 56 && b.pos < b.n) {
        b.buf |= (uint64_t)b.in[b.pos++] << b.cnt;
        b.cnt += 8;
    }
}
__device__ __forceinline__ uint32_t b5_bits(b5_bits_t& b, int n) { /* n <= 16 */
    if (b.cnt < n) {
        b5_refill(b);
        if (b.cnt < n) {
            b.bad = true;
            b.cnt = 64; /* feed zeros, the caller stops at the next check */
        }
    }
    const uint32_t v = (uint32_t)(b.buf & ((1ull << n) - 1ull));
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

/* canonical Huffman decode one bit at a time (codes longer than the direct table, and the code-length alphabet) */
__device__ __forceinline__ int b5_decode_slow(b5_bits_t& b, const uint16_t* cnt, const uint16_t* sym) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)b5_bits(b, 1);
        const int c = cnt[len];
        if (code - c < first) return sym[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    b.bad = true;
    return 0;
}

/* canonical code from the lengths len[0..n): cnt / sym for the slow path and, when tab != NULL, the direct table of
 * `bits` index bits (indexed by the next bits of the stream, which arrive least-significant first). Returns false on
 * an over-subscribed set of lengths. */
__device__ __forceinline__ bool b5_build(const uint8_t* len, int n, uint16_t* cnt, uint16_t* sym, uint16_t* tab, int bits) {
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int i = 0; i < n; i++) cnt[len[i]]++;
    cnt[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= cnt[l];
        if (left < 0) return false;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + cnt[l];
    for (int i = 0; i < n; i++)
        if (len[i]) sym[offs[len[i]]++] = (uint16_t)i;
    if (tab) {
        for (int i = 0; i < (1 << bits); i++) tab[i] = 0;
        int code = 0, idx = 0;
        for (int l = 1; l <= bits; l++) {
            for (int k = 0; k < cnt[l]; k++, code++, idx++) {
                /* the code, most significant bit first, reversed into stream order */
                uint32_t rev = __brev((uint32_t)code) >> (32 - l);
                const uint16_t e = (uint16_t)((sym[idx] << 4) | l);
                for (uint32_t j = rev; j < (1u << bits); j += (1u << l)) tab[j] = e;
            }
            code <<= 1;
        }
    }
    return true;
}

__device__ __forceinline__ int b5_decode(b5_bits_t& b, const uint16_t* tab, int bits, const uint16_t* cnt, const uint16_t* sym) {
    if (b.cnt < 15) b5_refill(b);
    const uint16_t e = tab[b.buf & ((1u << bits) - 1u)];
    if (e) {
        const int l = e & 15;
        if (b.cnt < l) {
            b.bad = true;
            return 0;
        }
        b.buf >>= l;
        b.cnt -= l;
        return e >> 4;
    }
    return b5_decode_slow(b, cnt, sym);
}

/* Lane 0 of a warp inflates one zlib stream (slow5lib: inflate() on the whole record, slow5_press.c:921-1010). */
__global__ void __launch_bounds__(32 * B5_INFLATE_WARPS)
abea_inflate_kernel(const abea_b5rec_t* __restrict__ recs, int32_t n, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                    int32_t* __restrict__ out_len, int32_t* __restrict__ status) {
#ifdef ABEA_SIMT_EMU
    unsigned char* dyn = (unsigned char*)simt::g_dynsmem;
#else
    extern __shared__ __align__(16) unsigned char b5_dyn_smem[];
    unsigned char* dyn = b5_dyn_smem;
#endif
    const int32_t r = (int32_t)(blockIdx.x * B5_INFLATE_WARPS + (threadIdx.x >> 5));
    if (r >= n || (threadIdx.x & 31) != 0) return;
    b5_tables_t& T = ((b5_tables_t*)dyn)[threadIdx.x >> 5];
    const abea_b5rec_t rec = recs[r];
    uint8_t* o = out + rec.out_off;
    const int64_t cap = rec.out_cap;
    int64_t op = 0;
    b5_bits_t b;
    b.in = in + rec.in_off;
    b.n = rec.in_len;
    b.pos = 0;
    b.buf = 0;
    b.cnt = 0;
    b.bad = false;
    int st = B5_OK;
    /* zlib wrapper (RFC 1950): CMF, FLG; no preset dictionary; the Adler-32 trailer is not verified */
    const uint32_t cmf = b5_bits(b, 8), flg = b5_bits(b, 8);
    if ((cmf & 15u) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) st = B5_ERR_DATA;
    const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint8_t clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    bool last = false;
    while (st == B5_OK && !last) {
        last = b5_bits(b, 1) != 0;
        const uint32_t type = b5_bits(b, 2);
        if (b.bad) { st = B5_ERR_DATA; break; }
        if (type == 0) { /* stored */
            const int drop = b.cnt & 7;
            b.buf >>= drop;
            b.cnt -= drop;
            const uint32_t len = b5_bits(b, 16), nlen = b5_bits(b, 16);
            if (b.bad || (len ^ 0xffffu) != nlen) { st = B5_ERR_DATA; break; }
            if (op + len > cap) { st = B5_ERR_OVERFLOW; break; }
            for (uint32_t i = 0; i < len; i++) o[op++] = (uint8_t)b5_bits(b, 8);
            if (b.bad) { st = B5_ERR_DATA; break; }
            continue;
        }
        if (type == 3) { st = B5_ERR_DATA; break; }
        int nlen_codes, ndist_codes;
        if (type == 1) { /* fixed code (RFC 1951 3.2.6) */
            for (int i = 0; i < 144; i++) T.len[i] = 8;
            for (int i = 144; i < 256; i++) T.len[i] = 9;
            for (int i = 256; i < 280; i++) T.len[i] = 7;
            for (int i = 280; i < 288; i++) T.len[i] = 8;
            for (int i = 0; i < 30; i++) T.len[288 + i] = 5;
            nlen_codes = 288;
            ndist_codes = 30;
        } else { /* dynamic code (3.2.7) */
            nlen_codes = (int)b5_bits(b, 5) + 257;
            ndist_codes = (int)b5_bits(b, 5) + 1;
            const int ncl = (int)b5_bits(b, 4) + 4;
            if (b.bad || nlen_codes > 286 || ndist_codes > 30) { st = B5_ERR_DATA; break; }
            uint8_t cl[19];
            for (int i = 0; i < 19; i++) cl[i] = 0;
            for (int i = 0; i < ncl; i++) cl[clorder[i]] = (uint8_t)b5_bits(b, 3);
            /* the code-length code: decoded bit by bit with the distance arrays as scratch (a few hundred symbols) */
            if (!b5_build(cl, 19, T.dcnt, T.dsym, nullptr, 0)) { st = B5_ERR_DATA; break; }
            int i = 0;
            while (i < nlen_codes + ndist_codes && !b.bad) {
                const int s = b5_decode_slow(b, T.dcnt, T.dsym);
                if (s < 16) {
                    T.len[i++] = (uint8_t)s;
                } else {
                    int prev = 0, rep;
                    if (s == 16) {
                        if (i == 0) { b.bad = true; break; }
                        prev = T.len[i - 1];
                        rep = 3 + (int)b5_bits(b, 2);
                    } else if (s == 17) {
                        rep = 3 + (int)b5_bits(b, 3);
                    } else {
                        rep = 11 + (int)b5_bits(b, 7);
                    }
                    if (i + rep > nlen_codes + ndist_codes) { b.bad = true; break; }
                    while (rep--) T.len[i++] = (uint8_t)prev;
                }
            }
            if (b.bad || T.len[256] == 0) { st = B5_ERR_DATA; break; }
            /* distance lengths follow the literal/length ones: move them to a fixed place */
            uint8_t dl[32];
            for (int j = 0; j < ndist_codes; j++) dl[j] = T.len[nlen_codes + j];
            for (int j = 0; j < ndist_codes; j++) T.len[288 + j] = dl[j];
        }
        if (!b5_build(T.len, nlen_codes, T.lcnt, T.lsym, T.lit, B5_LIT_BITS)) { st = B5_ERR_DATA; break; }
        /* an incomplete distance code is legal (a single distance code) */
        b5_build(T.len + 288, ndist_codes, T.dcnt, T.dsym, T.dist, B5_DIST_BITS);
        for (;;) {
            int s = b5_decode(b, T.lit, B5_LIT_BITS, T.lcnt, T.lsym);
            if (b.bad) { st = B5_ERR_DATA; break; }
            if (s < 256) {
                if (op >= cap) { st = B5_ERR_OVERFLOW; break; }
                o[op++] = (uint8_t)s;
            } else if (s == 256) {
                break;
            } else {
                s -= 257;
                if (s >= 29) { st = B5_ERR_DATA; break; }
                const int len = lbase[s] + (int)b5_bits(b, lext[s]);
                const int ds = b5_decode(b, T.dist, B5_DIST_BITS, T.dcnt, T.dsym);
                if (b.bad || ds >= 30) { st = B5_ERR_DATA; break; }
                const int64_t d = dbase[ds] + (int64_t)b5_bits(b, dext[ds]);
                if (b.bad || d > op) { st = B5_ERR_DATA; break; }
                if (op + len > cap) { st = B5_ERR_OVERFLOW; break; }
                for (int i = 0; i < len; i++, op++) o[op] = o[op - d];
            }
        }
    }
    out_len[r] = (int32_t)op;
    status[r] = st;
}

__device__ __forceinline__ uint64_t b5_le(const uint8_t* p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

/* The fields of a decompressed record up to the signal (slow5lib/src/slow5.c:2840-2930): uint16 length + read_id,
 * uint32 read_group, double digitisation, offset, range, sampling_rate, uint64 len_raw_signal — the number of samples,
 * or with signal compression the number of BYTES of the compressed signal, whose first four are the sample count
 * (slow5_press.c:1118-1121). */
__global__ void abea_blow5_parse_kernel(const abea_b5rec_t* __restrict__ recs, int32_t n, const uint8_t* __restrict__ data,
                                        const int32_t* __restrict__ rec_len, const int32_t* __restrict__ rec_status,
                                        int32_t signal_method, abea_b5hdr_t* __restrict__ hdr) {
    const int32_t r = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
    if (r >= n) return;
    abea_b5hdr_t h;
    h.digitisation = h.offset = h.range = h.sampling_rate = 0.0;
    h.sig_off = 0;
    h.sig_bytes = 0;
    h.n_samples = 0;
    h.status = rec_status ? rec_status[r] : B5_OK;
    const int64_t len = rec_len[r];
    const uint8_t* p = data + recs[r].out_off;
    if (h.status == B5_OK) {
        if (len < 2) {
            h.status = B5_ERR_DATA;
        } else {
            int64_t q = 2 + (int64_t)b5_le(p, 2);
            if (q + 4 + 32 + 8 > len) {
                h.status = B5_ERR_DATA;
            } else {
                q += 4; /* read_group */
                h.digitisation = __longlong_as_double((long long)b5_le(p + q, 8));
                h.offset = __longlong_as_double((long long)b5_le(p + q + 8, 8));
                h.range = __longlong_as_double((long long)b5_le(p + q + 16, 8));
                h.sampling_rate = __longlong_as_double((long long)b5_le(p + q + 24, 8));
                q += 32;
                const uint64_t lrs = b5_le(p + q, 8);
                q += 8;
                h.sig_off = q;
                if (signal_method == B5_SIG_NONE) {
                    h.sig_bytes = (int64_t)(2 * lrs);
                    h.n_samples = (int32_t)lrs;
                    if (lrs > 0x3fffffffull || q + h.sig_bytes > len) h.status = B5_ERR_DATA;
                } else if (signal_method == B5_SIG_EX_ZD) {
                    /* ex-zd v0 (slow5_press.c:1721-1842): u8 version, u64 samples, u8 shift, then ex_zd_press_16's stream */
                    h.sig_bytes = (int64_t)lrs;
                    if (lrs < 16 || lrs > 0x7fffffffull || q + h.sig_bytes > len || p[q] != 0) {
                        h.status = B5_ERR_DATA;
                    } else {
                        const uint64_t cnt = b5_le(p + q + 1, 8);
                        h.n_samples = (int32_t)cnt;
                        if (cnt < 1 || cnt > 0x3fffffffull || p[q + 9] > 5) h.status = B5_ERR_DATA;
                    }
                } else {
                    h.sig_bytes = (int64_t)lrs;
                    if (lrs < 4 || lrs > 0x7fffffffull || q + h.sig_bytes > len) {
                        h.status = B5_ERR_DATA;
                    } else {
                        const uint64_t cnt = b5_le(p + q, 4);
                        h.n_samples = (int32_t)cnt;
                        if (cnt > 0x3fffffffull || 4 + (cnt + 3) / 4 > lrs) h.status = B5_ERR_DATA;
                    }
                }
            }
        }
        if (h.status != B5_OK) h.n_samples = 0;
    }
    hdr[r] = h;
}

/* streamvbyte (slow5lib's __slow5_streamvbyte_decode: (count+3)/4 key bytes, then 1..4 data bytes per value) by one warp
 * into out[0 .. count): tiles of 128 values, each lane one key byte = four values, a warp scan of the per-lane byte counts
 * gives where each lane's data starts. `len` = the stream's stored size; false when it does not hold exactly count values. */
__device__ __forceinline__ bool b5_svb_decode_warp(const uint8_t* __restrict__ in, int64_t len, int64_t count,
                                                   uint32_t* __restrict__ out, int lane) {
    const int64_t nkeys = (count + 3) / 4;
    if (nkeys > len) return false;
    const uint8_t* keys = in;
    const uint8_t* dat = in + nkeys;
    const int64_t dat_len = len - nkeys;
    int64_t dpos = 0;
    for (int64_t k0 = 0; k0 < nkeys; k0 += 32) {
        const int64_t ki = k0 + lane;
        const uint32_t key = ki < nkeys ? keys[ki] : 0u;
        int nv = 0;
        if (ki < nkeys) {
            const int64_t left = count - 4 * ki;
            nv = left >= 4 ? 4 : (int)left;
        }
        int bl[4], mybytes = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            bl[q] = (q < nv) ? (int)((key >> (2 * q)) & 3u) + 1 : 0;
            mybytes += bl[q];
        }
        int incl = mybytes;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(ABEA_FULL, incl, d);
            if (lane >= d) incl += v;
        }
        const int tile_bytes = __shfl_sync(ABEA_FULL, incl, 31);
        if (dpos + tile_bytes > dat_len) return false;
        int64_t my = dpos + (incl - mybytes);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t v = 0;
            for (int t = 0; t < bl[q]; t++) v |= (uint32_t)dat[my + t] << (8 * t);
            my += bl[q];
            if (q < nv) out[4 * ki + q] = v;
        }
        dpos += tile_bytes;
    }
    return dpos == dat_len;
}

/* ex-zd (slow5_press.c:1262-1842: ex_press / ex_depress over the zigzag deltas of the samples, shifted right by q when
 * every sample has q trailing zero bits) by one warp. Stream: u8 version = 0, u64 n, u8 q, u16 first zigzag delta, u32
 * number of exceptions (zigzag deltas above 255), their positions (strictly increasing, stored as streamvbyte of
 * position[i] - position[i-1] - 1) and their values minus 256 (streamvbyte) — both behind a u32 byte count, or inline as
 * two u32 when there is exactly one — and then one byte for every zigzag delta that is not an exception, in order.
 * `ex` = 2 * n words of scratch for the exception positions and values. */
__device__ __forceinline__ bool b5_exzd_decode_warp(const uint8_t* __restrict__ p, int64_t len, int32_t n, uint32_t* __restrict__ ex,
                                                    float* __restrict__ dst, int lane) {
    if (len < 16) return false;
    const int qs = (int)p[9];
    const uint32_t first = (uint32_t)b5_le(p + 10, 2);
    const int64_t nex = (int64_t)b5_le(p + 12, 4);
    const int64_t nz = (int64_t)n - 1; /* zigzag deltas after the first */
    if (nex > nz) return false;
    int64_t off = 16;
    uint32_t* ex_pos = ex;
    uint32_t* ex_val = ex + n;
    if (nex > 1) {
        if (off + 4 > len) return false;
        const int64_t pos_len = (int64_t)b5_le(p + off, 4);
        off += 4;
        if (off + pos_len + 4 > len) return false;
        if (!b5_svb_decode_warp(p + off, pos_len, nex, ex_pos, lane)) return false;
        off += pos_len;
        const int64_t val_len = (int64_t)b5_le(p + off, 4);
        off += 4;
        if (off + val_len > len) return false;
        if (!b5_svb_decode_warp(p + off, val_len, nex, ex_val, lane)) return false;
        off += val_len;
        __syncwarp();
        /* position[i] = sum of the stored differences up to i, plus i (undelta_inplace_increasing_u32) */
        uint32_t carry = 0;
        for (int64_t i0 = 0; i0 < nex; i0 += 32) {
            const int64_t i = i0 + lane;
            uint32_t v = i < nex ? ex_pos[i] + (i > 0 ? 1u : 0u) : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(ABEA_FULL, incl, d);
                if (lane >= d) incl += u;
            }
            if (i < nex) ex_pos[i] = carry + incl;
            carry += __shfl_sync(ABEA_FULL, incl, 31);
        }
        __syncwarp();
    } else if (nex == 1) {
        if (off + 8 > len) return false;
        if (lane == 0) {
            ex_pos[0] = (uint32_t)b5_le(p + off, 4);
            ex_val[0] = (uint32_t)b5_le(p + off + 4, 4);
        }
        off += 8;
        __syncwarp();
    }
    const uint8_t* bytes = p + off;
    if (len - off != nz - nex) return false; /* ex_depress consumes exactly one byte per non-exception */
    /* sample 0, then tiles of 32 deltas: lane t of a tile decodes delta i (the i-th value after the first), which is an
     * exception iff it is the r-th one's position, r = the number of exceptions before i (binary search); else it is byte
     * i - r. The running sum wraps at 16 bits like the reference's int16 arithmetic (unzigdelta_u16_16). */
    int32_t prev = (int32_t)(int16_t)((first >> 1) ^ (0u - (first & 1u)));
    if (lane == 0) dst[0] = (float)(int16_t)((uint32_t)prev << qs);
    bool bad = false;
    for (int64_t i0 = 0; i0 < nz; i0 += 32) {
        const int64_t i = i0 + lane;
        int32_t z = 0;
        if (i < nz) {
            int64_t lo = 0, hi = nex; /* first exception whose position is >= i */
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((int64_t)ex_pos[mid] < i) lo = mid + 1;
                else hi = mid;
            }
            uint32_t zd;
            if (lo < nex && (int64_t)ex_pos[lo] == i) {
                zd = ex_val[lo] + 256u;
                if (zd > 0xffffu) bad = true;
            } else {
                zd = bytes[i - lo];
            }
            z = (int32_t)(int16_t)((zd >> 1) ^ (0u - (zd & 1u)));
        }
        int32_t incl = z;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t u = __shfl_up_sync(ABEA_FULL, incl, d);
            if (lane >= d) incl += u;
        }
        if (i < nz) dst[i + 1] = (float)(int16_t)((uint32_t)(int16_t)(prev + incl) << qs);
        prev += __shfl_sync(ABEA_FULL, incl, 31);
    }
    /* the exceptions' positions must be strictly increasing and inside the signal */
    for (int64_t j = lane; j < nex; j += 32)
        if ((int64_t)ex_pos[j] >= nz || (j > 0 && ex_pos[j] <= ex_pos[j - 1])) bad = true;
    return !__any_sync(ABEA_FULL, bad);
}

/* The signal of one record by one warp, widened to float at raw[raw_off[r] ..] (src/f5cio.c:461). svb-zd: tiles of
 * 128 values — each lane takes one key byte = four values: 2-bit codes give their byte counts, a warp scan of the
 * per-lane byte counts gives where each lane's data starts, the values are zigzag-decoded deltas and a second warp
 * scan of the per-lane delta sums (int32, wrapping like the reference's running sum) turns them into samples. */
#define B5_SIG_WARPS 4
__global__ void __launch_bounds__(32 * B5_SIG_WARPS)
abea_blow5_signal_kernel(const abea_b5rec_t* __restrict__ recs, int32_t n, const uint8_t* __restrict__ data,
                         const abea_b5hdr_t* __restrict__ hdr, const int64_t* __restrict__ raw_off, int32_t signal_method,
                         float* __restrict__ raw, int32_t* __restrict__ sig_status, uint32_t* __restrict__ ex_scratch) {
    const int lane = threadIdx.x & 31;
    const int32_t r = (int32_t)(blockIdx.x * B5_SIG_WARPS + (threadIdx.x >> 5));
    if (r >= n) return;
    const abea_b5hdr_t h = hdr[r];
    if (h.status != B5_OK || h.n_samples <= 0) {
        if (lane == 0) sig_status[r] = h.status;
        return;
    }
    const uint8_t* p = data + recs[r].out_off + h.sig_off;
    float* dst = raw + raw_off[r];
    const int32_t cnt = h.n_samples;
    if (signal_method == B5_SIG_NONE) {
        for (int32_t j = lane; j < cnt; j += 32) {
            const int16_t v = (int16_t)(uint16_t)(p[2 * (int64_t)j] | ((uint32_t)p[2 * (int64_t)j + 1] << 8));
            dst[j] = (float)v;
        }
        if (lane == 0) sig_status[r] = B5_OK;
        return;
    }
    if (signal_method == B5_SIG_EX_ZD) { /* scratch: two words per sample of the batch, this record's at 2 * raw_off */
        const bool ok = ex_scratch != nullptr && b5_exzd_decode_warp(p, h.sig_bytes, cnt, ex_scratch + 2 * raw_off[r], dst, lane);
        if (lane == 0) sig_status[r] = ok ? B5_OK : B5_ERR_DATA;
        return;
    }
    const int64_t nkeys = ((int64_t)cnt + 3) / 4;
    const uint8_t* keys = p + 4;
    const uint8_t* dat = keys + nkeys;
    const int64_t dat_len = h.sig_bytes - 4 - nkeys;
    int64_t dpos = 0;   /* bytes of data consumed so far (warp-uniform) */
    int32_t prev = 0;   /* running sample value (warp-uniform) */
    bool bad = false;
    for (int64_t k0 = 0; k0 < nkeys; k0 += 32) {
        const int64_t ki = k0 + lane;
        const uint32_t key = ki < nkeys ? keys[ki] : 0u;
        int nv = 0; /* values of this lane in the tile */
        if (ki < nkeys) {
            const int64_t left = (int64_t)cnt - 4 * ki;
            nv = left >= 4 ? 4 : (int)left;
        }
        int bl[4], mybytes = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            bl[q] = (q < nv) ? (int)((key >> (2 * q)) & 3u) + 1 : 0;
            mybytes += bl[q];
        }
        int incl = mybytes;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(ABEA_FULL, incl, d);
            if (lane >= d) incl += v;
        }
        const int tile_bytes = __shfl_sync(ABEA_FULL, incl, 31);
        int64_t my = dpos + (incl - mybytes);
        if (dpos + tile_bytes > dat_len) {
            bad = true;
            break;
        }
        int32_t delta[4], dsum = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t v = 0;
            for (int t = 0; t < bl[q]; t++) v |= (uint32_t)dat[my + t] << (8 * t);
            my += bl[q];
            const int32_t z = (int32_t)(v >> 1) ^ -(int32_t)(v & 1u); /* streamvbyte_zigzag.c:_zigzag_decode_32 */
            dsum += z;
            delta[q] = dsum; /* inclusive within the lane */
        }
        int32_t sincl = dsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t v = __shfl_up_sync(ABEA_FULL, sincl, d);
            if (lane >= d) sincl += v;
        }
        const int32_t base = prev + (sincl - dsum);
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (q < nv) dst[4 * ki + q] = (float)(int16_t)(base + delta[q]);
        prev += __shfl_sync(ABEA_FULL, sincl, 31);
        dpos += tile_bytes;
    }
    if (!bad && dpos != dat_len) bad = true; /* slow5_press.c:1125-1131: the decoder must consume exactly the stored bytes */
    if (lane == 0) sig_status[r] = bad ? B5_ERR_DATA : B5_OK;
}

/* int16 ADC counts -> the float samples the event detection reads (the widening of src/f5cio.c:461), flat */
__global__ void abea_i16_to_f32_kernel(const int16_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (float)src[i];
}
