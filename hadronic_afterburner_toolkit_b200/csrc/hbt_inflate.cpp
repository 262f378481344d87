// gzip / deflate decoder of the file reader, see hbt_inflate.h.
//
// How it differs from zlib's inflate: a 64-bit bit buffer refilled without a branch (an unaligned 8-byte load
// OR-ed above the bits still held: 56-63 valid bits after every refill, enough for a whole length + distance pair),
// 11-bit / 8-bit primary decode tables whose entries carry the symbol, its extra-bit count and its base value, up
// to three literals per refill, matches copied in 8-byte words, and a decode loop that only stops at the end of a
// block, when the output piece is full or when the input buffer runs low — no per-symbol state machine.  The text
// this reader sees is a stream of literals (digits: ~3.4 bits each) with few matches, which is exactly the case
// where the per-symbol overhead decides.
//
// The gzip framing (header flags, CRC-32 and length trailer, concatenated members) follows RFC 1952; the CRC itself is
// zlib's crc32().  Every validity check zlib makes on a deflate stream is made here too (code sets over-subscribed or
// incomplete, repeat without a previous length, missing end-of-block code, distance beyond the output so far,
// invalid symbols, stored-block length complement, truncated input).
#include "hbt_inflate.h"

#include <zlib.h>  // crc32()

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int kLitBits = 11, kDistBits = 8, kClBits = 7;
constexpr size_t kWindow = 32768;
constexpr size_t kPiece = 1u << 18;   // output produced per decode call
constexpr size_t kOutSlack = 320;     // a match (<= 258) copied in 8-byte words may run past the piece
constexpr size_t kInBuf = 1u << 20;
constexpr size_t kInPad = 64;         // zero bytes behind the valid input: 8-byte loads never leave the buffer
constexpr size_t kInLow = 2048;       // refill the input when fewer bytes are left (a dynamic block header is < 700 bytes)

// K_LIT2: two literals whose codes together fit the primary index (value = first | second << 8, nbits = both codes)
enum : uint32_t { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_SUB = 3, K_BAD = 4, K_LIT2 = 5 };
inline uint32_t mk(uint32_t kind, uint32_t nbits, uint32_t extra, uint32_t val) { return nbits | (kind << 8) | (extra << 12) | (val << 16); }
inline uint32_t e_nbits(uint32_t e) { return e & 0xffu; }
inline uint32_t e_kind(uint32_t e) { return (e >> 8) & 0xfu; }
inline uint32_t e_extra(uint32_t e) { return (e >> 12) & 0xfu; }
inline uint32_t e_val(uint32_t e) { return e >> 16; }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// table entry of a symbol whose code (or the part of it that indexes a second-level table) is l bits long: the bits to
// consume are the code AND its extra bits (one shift on the decoder's critical path; the extra value is cut out of
// a copy of the bit buffer off that path), the `extra` field holds the code length
inline uint32_t finish(uint32_t sym_entry, uint32_t l) { return mk(e_kind(sym_entry), l + e_extra(sym_entry), l, e_val(sym_entry)); }

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

// what a symbol of each alphabet decodes to (entry without its code length)
inline uint32_t litlen_symbol(int s) {
    if (s < 256) return mk(K_LIT, 0, 0, static_cast<uint32_t>(s));
    if (s == 256) return mk(K_EOB, 0, 0, 0);
    if (s < 286) return mk(K_LEN, 0, kLenExtra[s - 257], kLenBase[s - 257]);
    return mk(K_BAD, 0, 0, 0);
}
inline uint32_t dist_symbol(int s) { return s < 30 ? mk(K_LEN, 0, kDistExtra[s], kDistBase[s]) : mk(K_BAD, 0, 0, 0); }
inline uint32_t cl_symbol(int s) { return mk(K_LIT, 0, 0, static_cast<uint32_t>(s)); }

// Canonical Huffman decode table: `primary` index bits, longer codes through second-level tables.  Returns false for
// a code set zlib rejects: over-subscribed, or incomplete unless it is a single code of length 1 (`allow_single`).
// An all-zero set gives a table whose every entry is K_BAD (legal for the distance alphabet of a literal-only block).
// gzip starts a new block every 16-32 thousand symbols (~16 KB of this reader's text), so a 55 MB file carries
// thousands of dynamic headers: no allocation in here, and nothing that is not proportional to the table size.
struct Table {
    static constexpr size_t kCap = 4096;  // 2^11 primary entries + second-level tables (zlib's bound for 11 / 15 bits: 2342)
    uint32_t e[kCap];
    size_t size = 0;
    const uint32_t *data() const { return e; }
    bool empty() const { return size == 0; }
};

template <typename SymFn>
bool build_table(const uint8_t *lens, int n, int primary, bool allow_single, SymFn sym, Table &table) {
    int count[16] = {0};
    for (int s = 0; s < n; s++) count[lens[s]]++;
    int maxlen = 15;
    while (maxlen > 0 && count[maxlen] == 0) maxlen--;
    const uint32_t np = 1u << primary, pmask = np - 1u;
    const uint32_t bad = mk(K_BAD, 1, 0, 0);
    table.size = np;
    if (maxlen == 0) {
        for (uint32_t k = 0; k < np; k++) table.e[k] = bad;
        return true;
    }
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;  // over-subscribed
    }
    if (left > 0 && !(allow_single && maxlen == 1)) return false;  // incomplete
    if (left > 0) for (uint32_t k = 0; k < np; k++) table.e[k] = bad;  // (a complete code fills every entry below)
    uint32_t next[16];
    uint32_t code = 0;
    count[0] = 0;
    for (int l = 1; l <= 15; l++) {
        code = (code + static_cast<uint32_t>(count[l - 1])) << 1;
        next[l] = code;
    }
    uint16_t rev[320];
    uint8_t sub_bits[1u << kLitBits];
    const bool deep = maxlen > primary;
    if (deep) std::memset(sub_bits, 0, np);
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        rev[s] = static_cast<uint16_t>(reverse_bits(next[l]++, l));
        if (l > primary) {
            uint8_t &b = sub_bits[rev[s] & pmask];
            if (l - primary > b) b = static_cast<uint8_t>(l - primary);
        }
    }
    if (deep) {
        for (uint32_t p = 0; p <= pmask; p++) {
            if (!sub_bits[p]) continue;
            const size_t off = table.size, len = static_cast<size_t>(1) << sub_bits[p];
            if (off + len > Table::kCap) return false;
            for (size_t k = 0; k < len; k++) table.e[off + k] = bad;
            table.size = off + len;
            table.e[p] = mk(K_SUB, static_cast<uint32_t>(primary), sub_bits[p], static_cast<uint32_t>(off));
        }
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        if (l <= primary) {
            const uint32_t e = finish(sym(s), static_cast<uint32_t>(l));
            for (uint32_t k = rev[s]; k <= pmask; k += 1u << l) table.e[k] = e;
        } else {
            const uint32_t p = rev[s] & pmask;
            const uint32_t off = e_val(table.e[p]), sb = sub_bits[p];
            const uint32_t e = finish(sym(s), static_cast<uint32_t>(l - primary));
            for (uint32_t k = static_cast<uint32_t>(rev[s]) >> primary; k < (1u << sb); k += 1u << (l - primary)) table.e[off + k] = e;
        }
    }
    return true;
}

// Literal pairs: where the primary index holds a literal's code AND the whole code of the literal that follows it,
// one lookup yields both (the reader's text is ~3.4 bits per character: most lookups then produce two bytes).
void add_literal_pairs(Table &table, int primary) {
    const uint32_t n = 1u << primary;
    uint32_t first[1u << kLitBits];
    std::memcpy(first, table.e, n * sizeof(uint32_t));
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t a = first[i];
        if (e_kind(a) != K_LIT) continue;
        const uint32_t na = e_nbits(a);
        const uint32_t b = first[i >> na];  // the index bits above the first code, zero-extended: valid for a code that fits them
        if (e_kind(b) != K_LIT || na + e_nbits(b) > static_cast<uint32_t>(primary)) continue;
        table.e[i] = mk(K_LIT2, na + e_nbits(b), 0, e_val(a) | (e_val(b) << 8));
    }
}

inline uint64_t load64(const uint8_t *p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return v;  // little endian host (x86-64, aarch64)
}

}  // namespace

struct HbtGz {
    FILE *f = nullptr;
    std::string err;
    // input
    std::vector<uint8_t> inbuf;
    const uint8_t *in = nullptr, *in_end = nullptr;
    bool eof = false;
    uint64_t bitbuf = 0;
    int bitcnt = 0;
    // output: [kWindow bytes of history][piece + slack]
    std::vector<uint8_t> out;
    size_t out_pos = kWindow;   // next byte to write
    size_t handed = kWindow;    // next byte to hand to the caller
    long long member_origin = kWindow;  // position, in the coordinates of the current piece, of the member's first byte (<= 0 once it is history)
    uint32_t crc = 0;
    size_t crc_from = kWindow;  // first byte of the piece not yet folded into the CRC
    // decoder state
    enum State { S_START, S_PLAIN, S_GZ_HEADER, S_BLOCK, S_STORED, S_HUFF, S_TRAILER, S_DONE } state = S_START;
    bool final_block = false;
    uint32_t stored_left = 0;
    Table lit, dist, fixed_lit, fixed_dist, cl;
    const uint32_t *LT = nullptr, *DT = nullptr;

    bool fail(const char *m) { err = m; return false; }

    // keep at least kInLow bytes ahead (or everything up to the end of the file)
    bool fill_input() {
        if (eof || static_cast<size_t>(in_end - in) >= kInLow) return true;
        const size_t have = static_cast<size_t>(in_end - in);
        std::memmove(inbuf.data(), in, have);
        const size_t got = std::fread(inbuf.data() + have, 1, kInBuf - have, f);
        if (got < kInBuf - have) {
            if (std::ferror(f)) return fail("read error");
            eof = true;
        }
        in = inbuf.data();
        in_end = in + have + got;
        std::memset(inbuf.data() + have + got, 0, kInPad);
        return true;
    }

    // ---- bit reader for the headers (the symbol loop keeps its own copies in registers) ----
    void refill() {
        bitbuf |= load64(in) << bitcnt;
        const int adv = (63 - bitcnt) >> 3;
        in += adv;
        bitcnt += adv * 8;
    }
    uint32_t bits(int n) {  // n <= 32
        if (bitcnt < n) refill();
        const uint32_t v = static_cast<uint32_t>(bitbuf & ((1ull << n) - 1ull));
        bitbuf >>= n;
        bitcnt -= n;
        return v;
    }
    // drop the rest of the current byte and give the whole bytes still in the bit buffer back to the input
    void to_bytes() {
        in -= bitcnt >> 3;
        bitbuf = 0;
        bitcnt = 0;
    }
    // true when the consumed position lies beyond the end of the input (a truncated stream decodes zero padding)
    bool overrun() const { return eof && in - (bitcnt >> 3) > in_end; }

    bool gz_header() {
        // byte aligned here
        if (in_end - in < 10) return fail("truncated gzip header");
        if (in[0] != 0x1f || in[1] != 0x8b) return fail("not a gzip member");
        if (in[2] != 8) return fail("unknown compression method");
        const int flg = in[3];
        if (flg & 0xe0) return fail("unknown gzip header flags");
        in += 10;
        if (flg & 4) {  // FEXTRA
            if (in_end - in < 2) return fail("truncated gzip header");
            size_t xlen = in[0] | (in[1] << 8);
            in += 2;
            while (xlen) {
                if (!fill_input()) return false;
                if (in == in_end) return fail("truncated gzip header");
                const size_t k = std::min<size_t>(xlen, static_cast<size_t>(in_end - in));
                in += k;
                xlen -= k;
            }
        }
        for (int pass = 0; pass < 2; pass++) {  // FNAME, FCOMMENT: zero-terminated
            if (!(flg & (pass ? 16 : 8))) continue;
            for (;;) {
                if (!fill_input()) return false;
                if (in == in_end) return fail("truncated gzip header");
                if (*in++ == 0) break;
            }
        }
        if (flg & 2) {  // FHCRC
            if (!fill_input()) return false;
            if (in_end - in < 2) return fail("truncated gzip header");
            in += 2;
        }
        crc = static_cast<uint32_t>(crc32(0L, Z_NULL, 0));
        crc_from = out_pos;
        member_origin = static_cast<long long>(out_pos);  // a member's distances may not reach back into the previous one
        bitbuf = 0;
        bitcnt = 0;
        return true;
    }

    bool block_header() {
        final_block = bits(1) != 0;
        const uint32_t type = bits(2);
        if (type == 0) {
            // stored: skip to the byte boundary, LEN, NLEN
            const int drop = bitcnt & 7;
            bitbuf >>= drop;
            bitcnt -= drop;
            to_bytes();
            if (in_end - in < 4) return fail("truncated stored block");
            const uint32_t len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
            if ((len ^ 0xffffu) != nlen) return fail("invalid stored block lengths");
            in += 4;
            stored_left = len;
            state = S_STORED;
            return true;
        }
        if (type == 1) {
            if (fixed_lit.empty()) {
                uint8_t l[288];
                for (int s = 0; s < 288; s++) l[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
                build_table(l, 288, kLitBits, false, litlen_symbol, fixed_lit);
                add_literal_pairs(fixed_lit, kLitBits);
                uint8_t d[32];
                for (int s = 0; s < 32; s++) d[s] = 5;
                build_table(d, 32, kDistBits, false, dist_symbol, fixed_dist);
            }
            LT = fixed_lit.data();
            DT = fixed_dist.data();
            state = S_HUFF;
            return true;
        }
        if (type == 3) return fail("invalid block type");
        const int nlen = static_cast<int>(bits(5)) + 257, ndist = static_cast<int>(bits(5)) + 1, ncode = static_cast<int>(bits(4)) + 4;
        if (nlen > 286 || ndist > 30) return fail("too many length or distance symbols");
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl_lens[19] = {0};
        for (int k = 0; k < ncode; k++) cl_lens[order[k]] = static_cast<uint8_t>(bits(3));
        if (!build_table(cl_lens, 19, kClBits, false, cl_symbol, cl)) return fail("invalid code lengths set");
        uint8_t lens[320];
        int have = 0;
        while (have < nlen + ndist) {
            if (bitcnt < 15) refill();
            const uint32_t e = cl.e[bitbuf & ((1u << kClBits) - 1u)];
            if (e_kind(e) != K_LIT) return fail("invalid code lengths set");
            bitbuf >>= e_nbits(e);
            bitcnt -= static_cast<int>(e_nbits(e));
            const uint32_t s = e_val(e);
            if (s < 16) {
                lens[have++] = static_cast<uint8_t>(s);
                continue;
            }
            uint8_t v = 0;
            int rep;
            if (s == 16) {
                if (have == 0) return fail("invalid bit length repeat");
                v = lens[have - 1];
                rep = 3 + static_cast<int>(bits(2));
            } else if (s == 17) {
                rep = 3 + static_cast<int>(bits(3));
            } else {
                rep = 11 + static_cast<int>(bits(7));
            }
            if (have + rep > nlen + ndist) return fail("invalid bit length repeat");
            while (rep--) lens[have++] = v;
        }
        if (overrun()) return fail("unexpected end of file");
        if (lens[256] == 0) return fail("invalid code -- missing end-of-block");
        if (!build_table(lens, nlen, kLitBits, false, litlen_symbol, lit)) return fail("invalid literal/lengths set");
        add_literal_pairs(lit, kLitBits);
        if (!build_table(lens + nlen, ndist, kDistBits, true, dist_symbol, dist)) return fail("invalid distances set");
        LT = lit.data();
        DT = dist.data();
        state = S_HUFF;
        return true;
    }

    // the symbol loop: runs until the end of the block, a full piece or low input
    bool huff(size_t out_limit) {
        uint64_t bb = bitbuf;
        int bc = bitcnt;
        const uint8_t *ip = in;
        // at the end of the file the loop may run up to 8 bytes into the zero padding (bits loaded ahead of their use)
        const uint8_t *const ip_stop = eof ? in_end + 8 : in_end - kInLow / 2;
        uint8_t *const ob = out.data();
        size_t op = out_pos;
        const uint32_t *const lt = LT, *const dt = DT;
        const long long origin = member_origin;
        const char *bad = nullptr;
        bool eob = false;
#define HBT_REFILL()                                  \
    do {                                              \
        bb |= load64(ip) << bc;                       \
        const int adv_ = (63 - bc) >> 3;              \
        ip += adv_;                                   \
        bc += adv_ * 8;                               \
    } while (0)
        while (op < out_limit && ip < ip_stop) {
            // the lookup uses the bits already held, so that it does not wait for the refill's load (the refill only
            // adds bits above them)
            if (bc < kLitBits) HBT_REFILL();
            uint32_t e = lt[bb & ((1u << kLitBits) - 1u)];
            HBT_REFILL();
            // up to three lookups of literals (single or paired) on one refill: 3 x 11 bits <= 56
            int round = 0;
            for (;;) {
                const uint32_t k = e_kind(e);
                if (k == K_LIT2) {
                    bb >>= e_nbits(e); bc -= static_cast<int>(e_nbits(e));
                    const uint16_t two = static_cast<uint16_t>(e_val(e));
                    std::memcpy(ob + op, &two, 2);
                    op += 2;
                } else if (k == K_LIT) {
                    bb >>= e_nbits(e); bc -= static_cast<int>(e_nbits(e));
                    ob[op++] = static_cast<uint8_t>(e_val(e));
                } else {
                    break;
                }
                if (++round == 3) break;
                e = lt[bb & ((1u << kLitBits) - 1u)];
            }
            if (round) continue;  // (a non-literal that followed is looked up again after the next refill)
            if (e_kind(e) == K_SUB) {
                bb >>= kLitBits; bc -= kLitBits;
                e = lt[e_val(e) + (bb & ((1u << e_extra(e)) - 1u))];
            }
            uint64_t saved = bb;
            bb >>= e_nbits(e); bc -= static_cast<int>(e_nbits(e));  // the code and its extra bits
            if (e_kind(e) == K_LIT) {
                ob[op++] = static_cast<uint8_t>(e_val(e));
                continue;
            }
            if (e_kind(e) == K_EOB) { eob = true; break; }
            if (e_kind(e) != K_LEN) { bad = "invalid literal/length code"; break; }
            uint32_t len = e_val(e) + static_cast<uint32_t>((saved & ((1ull << e_nbits(e)) - 1ull)) >> e_extra(e));
            // <= 20 bits used since the refill; a distance code with its extra bits is <= 28 more
            e = dt[bb & ((1u << kDistBits) - 1u)];
            if (e_kind(e) == K_SUB) {
                bb >>= kDistBits; bc -= kDistBits;
                e = dt[e_val(e) + (bb & ((1u << e_extra(e)) - 1u))];
            }
            saved = bb;
            bb >>= e_nbits(e); bc -= static_cast<int>(e_nbits(e));
            if (e_kind(e) != K_LEN) { bad = "invalid distance code"; break; }
            const uint32_t d = e_val(e) + static_cast<uint32_t>((saved & ((1ull << e_nbits(e)) - 1ull)) >> e_extra(e));
            if (static_cast<long long>(d) > static_cast<long long>(op) - origin) { bad = "invalid distance too far back"; break; }
            // (d <= 32768 by the alphabet; the history in front of the piece holds the last 32768 bytes)
            uint8_t *dst = ob + op;
            const uint8_t *src = dst - d;
            op += len;
            if (d >= 8) {
                uint8_t *const end = dst + len;
                do {
                    std::memcpy(dst, src, 8);
                    dst += 8; src += 8;
                } while (dst < end);
            } else if (d == 1) {
                std::memset(dst, *src, len);
            } else {
                while (len--) *dst++ = *src++;
            }
        }
#undef HBT_REFILL
        bitbuf = bb;
        bitcnt = bc;
        in = ip;
        out_pos = op;
        if (bad) return fail(bad);
        if (overrun()) return fail("unexpected end of file");
        if (eob) state = final_block ? S_TRAILER : S_BLOCK;
        return true;
    }

    void fold_crc() {
        if (out_pos > crc_from) crc = static_cast<uint32_t>(crc32(crc, out.data() + crc_from, static_cast<uInt>(out_pos - crc_from)));
        crc_from = out_pos;
    }

    // decodes until a piece is full or the file ends; the piece is out[kWindow .. out_pos)
    bool produce() {
        // keep the last kWindow bytes as history in front of the new piece
        const size_t made = out_pos - kWindow;
        if (made) {
            if (made >= kWindow) std::memcpy(out.data(), out.data() + out_pos - kWindow, kWindow);
            else std::memmove(out.data(), out.data() + made, kWindow);
            member_origin -= static_cast<long long>(made);
        }
        out_pos = handed = crc_from = kWindow;
        const size_t limit = kWindow + kPiece;
        while (out_pos < limit && state != S_DONE) {
            if (!fill_input()) return false;
            switch (state) {
                case S_START:
                    state = (in_end - in >= 2 && in[0] == 0x1f && in[1] == 0x8b) ? S_GZ_HEADER : S_PLAIN;
                    break;
                case S_PLAIN: {
                    const size_t k = std::min<size_t>(limit - out_pos, static_cast<size_t>(in_end - in));
                    std::memcpy(out.data() + out_pos, in, k);
                    in += k;
                    out_pos += k;
                    if (k == 0 && eof) state = S_DONE;
                    break;
                }
                case S_GZ_HEADER:
                    if (!gz_header()) return false;
                    state = S_BLOCK;
                    break;
                case S_BLOCK:
                    if (!block_header()) return false;
                    break;
                case S_STORED: {
                    if (in > in_end) return fail("unexpected end of file");
                    const size_t k = std::min<size_t>(std::min<size_t>(stored_left, limit - out_pos), static_cast<size_t>(in_end - in));
                    std::memcpy(out.data() + out_pos, in, k);
                    in += k;
                    out_pos += k;
                    stored_left -= static_cast<uint32_t>(k);
                    if (stored_left == 0) state = final_block ? S_TRAILER : S_BLOCK;
                    else if (k == 0 && eof) return fail("unexpected end of file");
                    break;
                }
                case S_HUFF: {
                    const size_t before_out = out_pos;
                    const uint8_t *const before_in = in;
                    if (!huff(limit)) return false;
                    if (state == S_HUFF && out_pos == before_out && in == before_in && eof) return fail("unexpected end of file");
                    break;
                }
                case S_TRAILER: {
                    const int drop = bitcnt & 7;
                    bitbuf >>= drop;
                    bitcnt -= drop;
                    to_bytes();
                    if (in_end - in < 8) return fail("unexpected end of file");
                    fold_crc();
                    const uint32_t want_crc = in[0] | (in[1] << 8) | (in[2] << 16) | (static_cast<uint32_t>(in[3]) << 24);
                    const uint32_t want_len = in[4] | (in[5] << 8) | (in[6] << 16) | (static_cast<uint32_t>(in[7]) << 24);
                    in += 8;
                    if (want_crc != crc) return fail("incorrect data check");
                    if (want_len != static_cast<uint32_t>((static_cast<long long>(out_pos) - member_origin) & 0xffffffffll)) return fail("incorrect length check");
                    if (!fill_input()) return false;
                    // another member, or the end (bytes that are no gzip member are ignored, as gzread does after a complete one)
                    state = (in_end - in >= 2 && in[0] == 0x1f && in[1] == 0x8b) ? S_GZ_HEADER : S_DONE;
                    break;
                }
                case S_DONE:
                    break;
            }
        }
        if (state != S_PLAIN) fold_crc();
        return true;
    }
};

HbtGz *hbt_gz_open(const char *path) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return nullptr;
    HbtGz *g = new HbtGz;
    g->f = f;
    g->inbuf.resize(kInBuf + kInPad);
    g->in = g->in_end = g->inbuf.data();
    std::memset(g->inbuf.data(), 0, kInPad);
    g->out.resize(kWindow + kPiece + kOutSlack);
    return g;
}

long hbt_gz_read(HbtGz *g, char *buf, size_t n) {
    if (!g || !g->err.empty()) return -1;
    size_t done = 0;
    while (done < n) {
        if (g->handed == g->out_pos) {
            if (g->state == HbtGz::S_DONE) break;
            if (!g->produce()) return -1;
            if (g->handed == g->out_pos && g->state == HbtGz::S_DONE) break;
        }
        const size_t k = std::min(n - done, g->out_pos - g->handed);
        std::memcpy(buf + done, g->out.data() + g->handed, k);
        g->handed += k;
        done += k;
    }
    return static_cast<long>(done);
}

const char *hbt_gz_error(const HbtGz *g) { return g ? g->err.c_str() : "cannot open file"; }

void hbt_gz_close(HbtGz *g) {
    if (!g) return;
    if (g->f) std::fclose(g->f);
    delete g;
}
