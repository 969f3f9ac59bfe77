// inflate.cuh -- raw DEFLATE (RFC 1951) decoder for one BGZF block, usable on the device and on the host.
//
// BGZF (SAM spec 4.1) cuts a BAM file into independent deflate streams of at most 64 KiB, so a file is
// thousands of independent decode jobs: on the GPU a GROUP of lanes (an aligned part of a warp) inflates
// one block (k_bgzf_inflate, bam_gpu.cu).  All lanes of the group decode the symbol stream redundantly (identical state, broadcast loads), which
// keeps the control flow uniform; the work that can be split is split: an LZ77 match of `len` bytes is
// copied by the lanes in parallel -- byte k comes from dst[o - dist + k % dist], so even a run with
// dist = 1 has no serial dependency (one thread per block spent ~2 us per byte there, an L2 round trip
// each).  Huffman decoding is the canonical-code walk (one bit per step, count[] / symbol[] tables as in
// zlib's contrib/puff) -- no lookup tables, ~1.4 KB of per-lane local memory, no shared state.  Every loop
// consumes input or produces output and both are bounded, so malformed data ends in an error code,
// never in a hang.  On the host the same code runs with one "lane".
// (Measured on B200, 4 173 blocks of a 1 M-read BAM: 19.5 ms as written; with count[] packed into
// registers and the 15-step decode loop unrolled 46 ms at 111 registers, 30 ms capped at 64.)
//
// The same source compiles for the host: mcov_inflate_host() lets the CPU test suite check the decoder
// against zlib without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MCOV_HD __host__ __device__ __forceinline__
#else
#define MCOV_HD inline
#endif

namespace mcov {

#if defined(__CUDA_ARCH__)
#define MCOV_INF_SYNC() __syncwarp(gmask)   /* the cooperating lanes: a warp or an aligned part of one */
#define MCOV_INF_LDOUT(p) __ldcg(p)          /* output written by other lanes: read through L2 */
#else
#define MCOV_INF_SYNC() ((void)0)
#define MCOV_INF_LDOUT(p) (*(p))
#endif

enum InflateStatus {
  kInfOk = 0, kInfInputOverrun = 1, kInfOutputOverrun = 2, kInfBadBlockType = 3, kInfBadStored = 4,
  kInfBadCodeLengths = 5, kInfBadSymbol = 6, kInfBadDistance = 7, kInfShortOutput = 8
};

struct InfBits {
  const uint8_t* in;
  uint32_t n, p;
  uint64_t buf;
  int cnt;        // bits in buf
  int pad;        // how many of them (the most significant ones) are padding past the end of the input
  bool over;      // padding was CONSUMED: the stream is truncated
};

MCOV_HD void inf_need(InfBits& r, int need) {                  // make at least `need` (<= 32) bits available
  while (r.cnt < need) {
    uint64_t b = 0;
    if (r.p < r.n) b = r.in[r.p]; else r.pad += 8;
    ++r.p;
    r.buf |= b << r.cnt;
    r.cnt += 8;
  }
}
MCOV_HD uint32_t inf_peek(InfBits& r, int k) {
  inf_need(r, k);
  return (uint32_t)(r.buf & ((1ull << k) - 1ull));
}
MCOV_HD void inf_drop(InfBits& r, int k) {
  r.buf >>= k;
  r.cnt -= k;
  if (r.cnt < r.pad) r.over = true;
}
MCOV_HD uint32_t inf_bits(InfBits& r, int need) {
  const uint32_t v = inf_peek(r, need);
  inf_drop(r, need);
  return v;
}

struct InfHuff {
  int16_t count[16];
  int16_t* symbol;
};

// canonical Huffman code from code lengths; returns <0 over-subscribed, 0 complete, >0 incomplete
MCOV_HD int inf_construct(InfHuff& h, const int16_t* length, int n) {
  int16_t offs[16];
  for (int len = 0; len <= 15; ++len) h.count[len] = 0;
  for (int s = 0; s < n; ++s) h.count[length[s]]++;
  if (h.count[0] == n) return 0;
  int left = 1;
  for (int len = 1; len <= 15; ++len) {
    left <<= 1;
    left -= h.count[len];
    if (left < 0) return left;
  }
  offs[1] = 0;
  for (int len = 1; len < 15; ++len) offs[len + 1] = (int16_t)(offs[len] + h.count[len]);
  for (int s = 0; s < n; ++s)
    if (length[s] != 0) h.symbol[offs[length[s]]++] = (int16_t)s;
  return left;
}

MCOV_HD int inf_decode(InfBits& r, const InfHuff& h) {
  int code = 0, first = 0, index = 0;
  for (int len = 1; len <= 15; ++len) {
    code |= (int)inf_bits(r, 1);
    const int count = h.count[len];
    if (code - count < first) return h.symbol[index + (code - first)];
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

// First-level lookup tables: the next kLitBits (kDistBits) bits of the stream index an entry
// (symbol << 4 | code length) for every code of at most that many bits -- one lookup per symbol instead of
// one step per bit; longer (rare) codes have entry 0 and take the canonical walk above.  Built by the lanes
// together: the codes of one length are consecutive in symbol[] (canonical order), entry index = the
// bit-reversed code with every combination of the remaining high bits.
constexpr int kLitBits = 9, kDistBits = 6;
constexpr int kInfTabWords = (1 << kLitBits) + (1 << kDistBits);

MCOV_HD void inf_build_fast(const InfHuff& h, uint16_t* tab, int P, int lane, int nlanes, unsigned gmask) {
  (void)gmask;
  for (int i = lane; i < (1 << P); i += nlanes) tab[i] = 0;
  MCOV_INF_SYNC();
  int code = 0, base = 0;
  for (int len = 1; len <= P; ++len) {
    code = (code + h.count[len - 1] * (len > 1 ? 1 : 0)) << 1;   // first code of this length (count[0] does not enter)
    const int cnt = h.count[len];
    for (int t = lane; t < cnt; t += nlanes) {
      const int c = code + t;
      int rc = 0;
      for (int b = 0; b < len; ++b) rc |= ((c >> b) & 1) << (len - 1 - b);
      const uint16_t e = (uint16_t)((h.symbol[base + t] << 4) | len);
      for (int i = rc; i < (1 << P); i += (1 << len)) tab[i] = e;
    }
    base += cnt;
  }
  MCOV_INF_SYNC();
}

MCOV_HD int inf_decode_fast(InfBits& r, const InfHuff& h, const uint16_t* tab, int P) {
  const uint32_t e = tab[inf_peek(r, P)];
  if (e & 15u) { inf_drop(r, (int)(e & 15u)); return (int)(e >> 4); }
  return inf_decode(r, h);
}

// length / distance base values and extra bits of RFC 1951 3.2.5 in closed form (no tables in local memory)
MCOV_HD int inf_len_extra(int s) { return (s < 8 || s == 28) ? 0 : ((s - 4) >> 2); }
MCOV_HD int inf_len_base(int s) { return s < 8 ? 3 + s : (s == 28 ? 258 : 3 + ((4 + (s & 3)) << ((s - 4) >> 2))); }
MCOV_HD int inf_dist_extra(int s) { return s < 4 ? 0 : ((s - 2) >> 1); }
MCOV_HD int inf_dist_base(int s) { return s < 4 ? 1 + s : 1 + ((2 + (s & 1)) << ((s - 2) >> 1)); }

// Inflate one raw deflate stream of clen bytes into exactly ulen bytes.  Returns an InflateStatus.
// Called by `nlanes` cooperating lanes (a warp, or 1 on the host) with identical arguments except `lane`;
// every lane returns the same status.
// `win` (optional, shared by the lanes): a circular window of wmask+1 bytes (a power of two) that mirrors the
// most recent output, so that match copies read shared memory instead of making a round trip to L2 for bytes
// the warp has just stored (measured: ~1.2 us per symbol without it, almost all of it that round trip).
// Matches that reach further back than the window read the output itself.
MCOV_HD int inflate_raw(const uint8_t* src, uint32_t clen, uint8_t* dst, uint32_t ulen, uint16_t* tabs /* kInfTabWords, shared by the lanes */,
                        int lane = 0, int nlanes = 1, uint8_t* win = nullptr, uint32_t wmask = 0, unsigned gmask = 0xffffffffu) {
  (void)gmask;
  InfBits r;
  r.in = src; r.n = clen; r.p = 0; r.buf = 0; r.cnt = 0; r.pad = 0; r.over = false;
  uint16_t* ltab = tabs;
  uint16_t* dtab = tabs + (1 << kLitBits);
  int16_t lengths[320];
  int16_t lensym[288], distsym[30];
  InfHuff lencode, distcode;
  lencode.symbol = lensym;
  distcode.symbol = distsym;
  uint32_t o = 0;
  int last;
  do {
    last = (int)inf_bits(r, 1);
    const int type = (int)inf_bits(r, 2);
    if (r.over) return kInfInputOverrun;
    if (type == 0) {                                        // stored
      const int drop = r.cnt & 7;
      (void)inf_bits(r, drop);
      const uint32_t len = inf_bits(r, 16), nlen = inf_bits(r, 16);
      if (r.over || (len ^ 0xffffu) != nlen) return kInfBadStored;
      if (o + len > ulen) return kInfOutputOverrun;
      for (uint32_t k = 0; k < len; ++k) {
        const uint8_t v = (uint8_t)inf_bits(r, 8);
        if (lane == 0) { dst[o] = v; if (win) win[o & wmask] = v; }
        ++o;
      }
      if (r.over) return kInfInputOverrun;
      continue;
    }
    if (type == 3) return kInfBadBlockType;
    if (type == 1) {                                        // fixed codes
      int s = 0;
      for (; s < 144; ++s) lengths[s] = 8;
      for (; s < 256; ++s) lengths[s] = 9;
      for (; s < 280; ++s) lengths[s] = 7;
      for (; s < 288; ++s) lengths[s] = 8;
      inf_construct(lencode, lengths, 288);
      for (s = 0; s < 30; ++s) lengths[s] = 5;
      inf_construct(distcode, lengths, 30);
    } else {                                                // dynamic codes
      const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      const int nlen = (int)inf_bits(r, 5) + 257, ndist = (int)inf_bits(r, 5) + 1, ncode = (int)inf_bits(r, 4) + 4;
      if (r.over) return kInfInputOverrun;
      if (nlen > 286 || ndist > 30) return kInfBadCodeLengths;
      int idx = 0;
      for (; idx < ncode; ++idx) lengths[order[idx]] = (int16_t)inf_bits(r, 3);
      for (; idx < 19; ++idx) lengths[order[idx]] = 0;
      if (inf_construct(lencode, lengths, 19) != 0) return kInfBadCodeLengths;
      idx = 0;
      while (idx < nlen + ndist) {
        int sym = inf_decode(r, lencode);
        if (sym < 0 || r.over) return kInfBadCodeLengths;
        if (sym < 16) {
          lengths[idx++] = (int16_t)sym;
        } else {
          int len = 0, rep;
          if (sym == 16) {
            if (idx == 0) return kInfBadCodeLengths;
            len = lengths[idx - 1];
            rep = 3 + (int)inf_bits(r, 2);
          } else if (sym == 17) {
            rep = 3 + (int)inf_bits(r, 3);
          } else {
            rep = 11 + (int)inf_bits(r, 7);
          }
          if (idx + rep > nlen + ndist) return kInfBadCodeLengths;
          while (rep--) lengths[idx++] = (int16_t)len;
        }
      }
      if (lengths[256] == 0) return kInfBadCodeLengths;
      int err = inf_construct(lencode, lengths, nlen);
      if (err < 0 || (err > 0 && nlen - lencode.count[0] != 1)) return kInfBadCodeLengths;
      err = inf_construct(distcode, lengths + nlen, ndist);
      if (err < 0 || (err > 0 && ndist - distcode.count[0] != 1)) return kInfBadCodeLengths;
    }
    inf_build_fast(lencode, ltab, kLitBits, lane, nlanes, gmask);
    inf_build_fast(distcode, dtab, kDistBits, lane, nlanes, gmask);
    // literal/length + distance symbols until end-of-block
    for (;;) {
      int sym = inf_decode_fast(r, lencode, ltab, kLitBits);
      if (sym < 0) return kInfBadSymbol;
      if (r.over) return kInfInputOverrun;
      if (sym < 256) {
        if (o >= ulen) return kInfOutputOverrun;
        if (lane == 0) { dst[o] = (uint8_t)sym; if (win) win[o & wmask] = (uint8_t)sym; }
        ++o;
      } else if (sym == 256) {
        break;
      } else {
        sym -= 257;
        if (sym >= 29) return kInfBadSymbol;
        const uint32_t len = (uint32_t)inf_len_base(sym) + inf_bits(r, inf_len_extra(sym));
        const int ds = inf_decode_fast(r, distcode, dtab, kDistBits);
        if (ds < 0 || ds >= 30) return kInfBadSymbol;
        const uint32_t dist = (uint32_t)inf_dist_base(ds) + inf_bits(r, inf_dist_extra(ds));
        if (r.over) return kInfInputOverrun;
        if (dist > o) return kInfBadDistance;
        if (o + len > ulen) return kInfOutputOverrun;
        // the match source [o - dist, o) is complete: byte k of the match repeats it with period dist
        MCOV_INF_SYNC();
        if (win && dist <= wmask + 1u - 258u) {
          // source [o-dist, o) and destination [o, o+len) lie within one window length of each other
          // (dist + len <= window), so their circular indices never collide
          const uint32_t so = o - dist;
          for (uint32_t k = (uint32_t)lane; k < len; k += (uint32_t)nlanes) {
            const uint8_t b = win[(so + (dist >= len ? k : k % dist)) & wmask];
            win[(o + k) & wmask] = b;
            dst[o + k] = b;
          }
        } else {
          const uint8_t* from = dst + (o - dist);
          for (uint32_t k = (uint32_t)lane; k < len; k += (uint32_t)nlanes) {
            const uint8_t b = MCOV_INF_LDOUT(from + (dist >= len ? k : k % dist));
            dst[o + k] = b;
            if (win) win[(o + k) & wmask] = b;
          }
        }
        MCOV_INF_SYNC();
        o += len;
      }
    }
  } while (!last);
  MCOV_INF_SYNC();
  return o == ulen ? kInfOk : kInfShortOutput;
}

// CRC-32 (IEEE 802.3, as in gzip) of n bytes, 4 bits per step from a 16-entry table
MCOV_HD uint32_t crc32_bytes(const uint8_t* p, uint32_t n) {
  const uint32_t t[16] = {0x00000000u, 0x1db71064u, 0x3b6e20c8u, 0x26d930acu, 0x76dc4190u, 0x6b6b51f4u, 0x4db26158u, 0x5005713cu,
                          0xedb88320u, 0xf00f9344u, 0xd6d6a3e8u, 0xcb61b38cu, 0x9b64c2b0u, 0x86d3d2d4u, 0xa00ae278u, 0xbdbdf21cu};
  uint32_t c = 0xffffffffu;
  for (uint32_t k = 0; k < n; ++k) {
    c ^= p[k];
    c = t[c & 15u] ^ (c >> 4);
    c = t[c & 15u] ^ (c >> 4);
  }
  return c ^ 0xffffffffu;
}


// ---- CRC-32 of a block by cooperating lanes -------------------------------------------------------
// CRC is linear over GF(2): crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B) for finalised CRCs (the
// identity behind zlib's crc32_combine).  Each lane takes a contiguous slice, the slice CRCs are folded
// left to right with ~32 polynomial multiplications -- instead of one lane walking all 64 KiB.
MCOV_HD uint32_t crc_multmodp(uint32_t a, uint32_t b) {          // a(x) * b(x) mod P, reflected bit order; a != 0
  uint32_t m = 1u << 31, p = 0;
  for (int it = 0; it < 32; ++it) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1u)) == 0) break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;
  }
  return p;
}

MCOV_HD uint32_t crc_x8n(uint32_t n) {                           // x^(8n) mod P
  uint32_t cur = 1u << 30;                                       // x^1
  cur = crc_multmodp(cur, cur); cur = crc_multmodp(cur, cur); cur = crc_multmodp(cur, cur);   // x^8
  uint32_t acc = 1u << 31;                                       // x^0
  while (n) {
    if (n & 1u) acc = crc_multmodp(cur, acc);
    cur = crc_multmodp(cur, cur);
    n >>= 1;
  }
  return acc;
}

// slice of lane `lane` out of `nlanes` over n bytes
MCOV_HD void crc_slice(uint32_t n, int lane, int nlanes, uint32_t& lo, uint32_t& hi) {
  const uint32_t per = (n + (uint32_t)nlanes - 1u) / (uint32_t)nlanes;
  lo = per * (uint32_t)lane; if (lo > n) lo = n;
  hi = lo + per; if (hi > n) hi = n;
}

// host form of the lane-sliced CRC (what the warp computes with shuffles in k_bgzf_inflate)
inline uint32_t crc32_sliced_host(const uint8_t* p, uint32_t n, int nlanes) {
  uint32_t crc = 0;
  for (int j = 0; j < nlanes; ++j) {
    uint32_t lo, hi;
    crc_slice(n, j, nlanes, lo, hi);
    const uint32_t cj = crc32_bytes(p + lo, hi - lo);
    crc = (j == 0) ? cj : ((hi > lo) ? (crc_multmodp(crc_x8n(hi - lo), crc) ^ cj) : crc);
  }
  return crc;
}

}  // namespace mcov
