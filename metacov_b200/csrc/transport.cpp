// transport.cpp -- host side of the transport block (include/metacov_b200.h: mcov_block_hdr, mcov_pack_block).
//
// What the native decoder hands to the GPU for a batch of coordinate-sorted reads: one contiguous buffer, as
// narrow as the data allows (the end-to-end rate is bound by the PCIe link).  The reference has no counterpart:
// it walks `bam1_t` records one by one on the host (metacov/scan.pyx:243-294).
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "metacov_b200.h"

namespace {

inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

struct CigKey { uint32_t n; uint32_t op[4]; };        // CIGARs of up to four ops are dictionary candidates
inline uint64_t cig_hash(const uint32_t* ops, uint32_t n) {
  uint64_t h = 1469598103934665603ull ^ n;
  for (uint32_t k = 0; k < n; ++k) { h ^= ops[k]; h *= 1099511628211ull; }
  return h;
}

constexpr int kTable = 1 << 14;                       // open-addressing table of candidate CIGARs

struct Cand { CigKey key; uint64_t count; bool used; };

template <typename F>
void par_for(int64_t n, int nt, F fn) {
  if (nt < 1) nt = 1;
  if (n < (1 << 16)) nt = 1;
  const int64_t per = (n + nt - 1) / nt;
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) { const int64_t a = t * per, b = std::min(n, a + per); if (a < b) th.emplace_back(fn, t, a, b); }
  fn(0, 0, std::min(n, per));
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int64_t mcov_block_bound(int64_t n, int64_t n_cigar, int32_t n_contigs) {
  if (n < 0 || n_cigar < 0 || n_contigs < 0) return -1;
  const size_t n1 = (size_t)std::max<int64_t>(n, 1);
  return (int64_t)(al16(sizeof(mcov_block_hdr)) + al16(((size_t)n_contigs + 1) * 8) + al16(n1) /*dpos*/ + 2 * al16(n1 * 4) /*exceptions*/ +
                   al16(n1) /*fc*/ + al16(1024) /*jt*/ + al16(n1 * 4) + al16(n1 * 2) + al16(n1) /*escapes*/ + al16(129 * 4) +
                   al16(128 * 4 * 4) + al16((size_t)n_cigar * 4 + 16) + al16(n1) /*mapq*/ + al16((n1 / MCOV_BLOCK_CHUNK + 2) * sizeof(mcov_block_chunk)) + 320);
}

int mcov_pack_block(int64_t n, int64_t n_carry, const int32_t* tid, const int32_t* pos, const uint16_t* flag,
                    const uint8_t* mapq, const uint32_t* cig_off, const uint32_t* cig, int32_t n_contigs,
                    void* out, int64_t cap, int64_t* bytes_out, int n_threads) {
  if (n < 0 || n_carry < 0 || n_carry > n || n_contigs <= 0 || !out || !bytes_out || (reinterpret_cast<uintptr_t>(out) & 15u)) return MCOV_ERR_ARG;
  if (n > 0 && (!tid || !pos || !flag || !cig_off)) return MCOV_ERR_ARG;
  if (n >= 0xFFFFFFF0ll) return MCOV_ERR_RANGE;
  const int64_t n_cig = n > 0 ? (int64_t)cig_off[n] : 0;
  if (cap < mcov_block_bound(n, n_cig, n_contigs)) return MCOV_ERR_ARG;
  if (n_threads <= 0) { unsigned hw = std::thread::hardware_concurrency(); n_threads = hw ? (int)std::min(hw, 32u) : 4; }
  try {
    char* base = static_cast<char*>(out);
    mcov_block_hdr h;
    std::memset(&h, 0, sizeof(h));
    h.magic = MCOV_BLOCK_MAGIC; h.version = 4; h.n = n; h.n_carry = n_carry; h.n_cigar = n_cig; h.n_contigs = n_contigs;
    h.last_tid = n > 0 ? tid[n - 1] : -1; h.last_pos = n > 0 ? pos[n - 1] : 0;
    h.has_mapq = mapq ? 1 : 0;
    size_t o = al16(sizeof(mcov_block_hdr));
    // ---- contig prefix (and the check that the batch is grouped by contig, unplaced reads last) ----
    h.off_crs = (uint32_t)o;
    int64_t* crs = reinterpret_cast<int64_t*>(base + o);
    o += al16(((size_t)n_contigs + 1) * 8);
    {
      int64_t i = 0;
      for (int32_t c = 0; c < n_contigs; ++c) {
        crs[c] = i;
        while (i < n && tid[i] == c) ++i;
      }
      crs[n_contigs] = i;
      for (int64_t k = i; k < n; ++k) if (tid[k] >= 0 && tid[k] < n_contigs) return MCOV_ERR_ARG;   // not grouped by contig
    }
    const size_t n1 = (size_t)std::max<int64_t>(n, 1);
    // ---- positions: u8 differences + exceptions ----
    std::vector<uint8_t> dpos_v(n1), fc_v(n1);                 // placed (wide or as nibbles) once both are known
    uint8_t* dpos = dpos_v.data();
    std::vector<std::vector<std::pair<uint32_t, int32_t>>> exc((size_t)n_threads);
    // ---- CIGAR dictionary: the 128 most frequent CIGARs of up to four ops (counted on a sample of the batch) ----
    std::vector<Cand> table(kTable);
    for (auto& c : table) { c.used = false; c.count = 0; }
    const int64_t step = std::max<int64_t>(1, n / 200000);
    for (int64_t i = 0; i < n; i += step) {
      const uint32_t nc = cig_off[i + 1] - cig_off[i];
      if (nc == 0 || nc > 4) continue;
      const uint32_t* ops = cig + cig_off[i];
      uint64_t hsh = cig_hash(ops, nc);
      for (int probe = 0; probe < 64; ++probe) {
        Cand& c = table[(hsh + probe) & (kTable - 1)];
        if (!c.used) { c.used = true; c.key.n = nc; std::memcpy(c.key.op, ops, nc * 4); c.count = 1; break; }
        if (c.key.n == nc && std::memcmp(c.key.op, ops, nc * 4) == 0) { ++c.count; break; }
      }
    }
    std::vector<const Cand*> top;
    for (const auto& c : table) if (c.used) top.push_back(&c);
    std::sort(top.begin(), top.end(), [](const Cand* a, const Cand* b) { return a->count > b->count; });
    if (top.size() > 128) top.resize(128);
    // lookup table of the chosen entries (same hashing)
    std::vector<int16_t> chosen(kTable, -1);
    std::vector<CigKey> dict;
    for (const Cand* c : top) {
      const uint64_t hsh = cig_hash(c->key.op, c->key.n);
      for (int probe = 0; probe < kTable; ++probe) {
        int16_t& slot = chosen[(hsh + probe) & (kTable - 1)];
        if (slot < 0) { slot = (int16_t)dict.size(); break; }
      }
      dict.push_back(c->key);
    }
    auto dict_find = [&](const uint32_t* ops, uint32_t nc) -> int {
      if (nc == 0 || nc > 4 || dict.empty()) return -1;
      const uint64_t hsh = cig_hash(ops, nc);
      for (int probe = 0; probe < kTable; ++probe) {
        const int16_t slot = chosen[(hsh + probe) & (kTable - 1)];
        if (slot < 0) return -1;
        const CigKey& k = dict[(size_t)slot];
        if (k.n == nc && std::memcmp(k.op, ops, nc * 4) == 0) return slot;
      }
      return -1;
    };
    h.off_dict_off = (uint32_t)o;
    uint32_t* dict_off = reinterpret_cast<uint32_t*>(base + o);
    o += al16(129 * 4);
    h.off_dict_ops = (uint32_t)o;
    uint32_t* dict_ops = reinterpret_cast<uint32_t*>(base + o);
    o += al16(128 * 4 * 4);
    h.n_dict = (int32_t)dict.size();
    {
      uint32_t k = 0;
      for (size_t d = 0; d < dict.size(); ++d) { dict_off[d] = k; std::memcpy(dict_ops + k, dict[d].op, dict[d].n * 4); k += dict[d].n; }
      dict_off[dict.size()] = k;
      h.n_dictops = (int32_t)k;
    }
    // ---- per-read pass 1 (parallel): position differences, CIGAR classes; explicit op counts per range ----
    std::vector<uint8_t> cls((size_t)n1);
    std::vector<int64_t> xcount((size_t)n_threads + 1, 0);
    std::vector<int> too_long((size_t)n_threads, 0), wide_op((size_t)n_threads, 0);
    par_for(n, n_threads, [&](int t, int64_t a, int64_t b) {
      int64_t xc = 0;
      for (int64_t i = a; i < b; ++i) {
        // first read of a contig (or of the unplaced tail): difference to 0
        const bool first = i == 0 || tid[i] != tid[i - 1];
        const int64_t d = first ? (int64_t)pos[i] : (int64_t)pos[i] - (int64_t)pos[i - 1];
        if (d < 0 || d > 255) { dpos[i] = 0; exc[(size_t)t].emplace_back((uint32_t)i, (int32_t)d); }
        else dpos[i] = (uint8_t)d;
        const uint32_t nc = cig_off[i + 1] - cig_off[i];
        const int k = dict_find(cig + cig_off[i], nc);
        if (k >= 0) cls[(size_t)i] = (uint8_t)k;
        else if (nc <= 127) {
          cls[(size_t)i] = (uint8_t)(128 + nc); xc += nc;
          for (uint32_t q = 0; q < nc; ++q) if (cig[cig_off[i] + q] > 0xFFFFu) wide_op[(size_t)t] = 1;   // len >= 4096: no u16 form
        }
        else { too_long[(size_t)t] = 1; cls[(size_t)i] = 128; }
      }
      xcount[(size_t)t + 1] = xc;
    });
    for (int t = 0; t < n_threads; ++t) if (too_long[(size_t)t]) return MCOV_ERR_RANGE;
    h.xop_bytes = 2;
    for (int t = 0; t < n_threads; ++t) if (wide_op[(size_t)t]) h.xop_bytes = 4;
    // ---- joint table: the 255 most frequent (flag, class) pairs of a sample of the batch ----
    std::vector<uint32_t> jt;
    std::vector<int16_t> jslot(1u << 16, -1);                 // hash table over (flag << 8 | class), open addressing
    std::vector<uint32_t> jkey(1u << 16, 0xFFFFFFFFu);
    {
      std::vector<uint32_t> cnt(1u << 16, 0);
      auto slot_of = [&](uint32_t key) -> uint32_t {
        uint32_t hsh = (key * 2654435761u) >> 16;
        for (int probe = 0; probe < 65536; ++probe) {
          const uint32_t sl = (hsh + (uint32_t)probe) & 0xFFFFu;
          if (jkey[sl] == key || jkey[sl] == 0xFFFFFFFFu) return sl;
        }
        return 0xFFFFFFFFu;
      };
      size_t distinct = 0;
      for (int64_t i = 0; i < n; i += step) {
        const uint32_t key = ((uint32_t)flag[i] << 8) | cls[(size_t)i];
        if (distinct >= 60000) break;                          // (a pathological batch: enough candidates seen)
        const uint32_t sl = slot_of(key);
        if (sl == 0xFFFFFFFFu) break;
        if (jkey[sl] == 0xFFFFFFFFu) { jkey[sl] = key; ++distinct; }
        ++cnt[sl];
      }
      std::vector<uint32_t> order;
      for (uint32_t sl = 0; sl < (1u << 16); ++sl) if (jkey[sl] != 0xFFFFFFFFu) order.push_back(sl);
      std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cnt[x] > cnt[y]; });
      if (order.size() > 255) order.resize(255);
      for (uint32_t sl : order) { jslot[sl] = (int16_t)jt.size(); jt.push_back(jkey[sl]); }
    }
    auto jt_find = [&](uint32_t key) -> int {
      uint32_t hsh = (key * 2654435761u) >> 16;
      for (int probe = 0; probe < 65536; ++probe) {
        const uint32_t sl = (hsh + (uint32_t)probe) & 0xFFFFu;
        if (jkey[sl] == key) return jslot[sl];
        if (jkey[sl] == 0xFFFFFFFFu) return -1;
      }
      return -1;
    };
    uint8_t* fc = fc_v.data();
    h.off_jt = (uint32_t)o;
    h.n_jt = (int32_t)jt.size();
    std::memset(base + o, 0, 1024);
    std::memcpy(base + o, jt.data(), jt.size() * 4);           // entry: flag << 8 | class
    o += al16(1024);
    h.off_xops = (uint32_t)o;
    uint32_t* xops = reinterpret_cast<uint32_t*>(base + o);
    uint16_t* xops16 = reinterpret_cast<uint16_t*>(base + o);
    const bool narrow = h.xop_bytes == 2;
    for (int t = 0; t < n_threads; ++t) xcount[(size_t)t + 1] += xcount[(size_t)t];
    h.n_xops = xcount[(size_t)n_threads];
    // ---- per-read pass 2 (parallel): joint-table indices (escapes listed), explicit ops ----
    std::vector<std::vector<uint32_t>> esc((size_t)n_threads);
    par_for(n, n_threads, [&](int t, int64_t a, int64_t b) {
      int64_t w = xcount[(size_t)t];
      for (int64_t i = a; i < b; ++i) {
        const int k = jt_find(((uint32_t)flag[i] << 8) | cls[(size_t)i]);
        if (k >= 0) fc[i] = (uint8_t)k; else { fc[i] = 255; esc[(size_t)t].push_back((uint32_t)i); }
        if (cls[(size_t)i] >= 128) {
          const uint32_t nc = cls[(size_t)i] - 128u;
          if (narrow) for (uint32_t q = 0; q < nc; ++q) xops16[w + q] = (uint16_t)cig[cig_off[i] + q];
          else std::memcpy(xops + w, cig + cig_off[i], (size_t)nc * 4);
          w += nc;
        }
      }
    });
    o += al16((size_t)h.n_xops * (size_t)h.xop_bytes + 16);
    // ---- the two per-read bytes: wide (dpos[], fc[]) or as nibbles + side lists, whichever is smaller; per chunk of
    //      MCOV_BLOCK_CHUNK reads, where its entries of the side lists and of the explicit ops begin ----
    const int64_t n_chunks = (n + MCOV_BLOCK_CHUNK - 1) / MCOV_BLOCK_CHUNK;
    std::vector<uint32_t> cd((size_t)n_chunks + 1, 0), cf((size_t)n_chunks + 1, 0), cx((size_t)n_chunks + 1, 0);
    // (threads split the READS; a thread takes the chunks that begin in its range)
    auto chunks_of = [](int64_t a) { return (a + MCOV_BLOCK_CHUNK - 1) / MCOV_BLOCK_CHUNK; };
    par_for(n, n_threads, [&](int, int64_t ra, int64_t rb) {
      for (int64_t c = chunks_of(ra); c < chunks_of(rb); ++c) {
        uint32_t kd = 0, kf = 0, kx = 0;
        const int64_t e = std::min<int64_t>(n, (c + 1) * MCOV_BLOCK_CHUNK);
        for (int64_t i = c * MCOV_BLOCK_CHUNK; i < e; ++i) {
          kd += dpos[i] >= 15; kf += fc[i] >= 15;
          if (cls[(size_t)i] >= 128) kx += cls[(size_t)i] - 128u;
        }
        cd[(size_t)c + 1] = kd; cf[(size_t)c + 1] = kf; cx[(size_t)c + 1] = kx;
      }
    });
    for (int64_t c = 0; c < n_chunks; ++c) { cd[(size_t)c + 1] += cd[(size_t)c]; cf[(size_t)c + 1] += cf[(size_t)c]; cx[(size_t)c + 1] += cx[(size_t)c]; }
    {
      const size_t n_dq = cd[(size_t)n_chunks], n_fq = cf[(size_t)n_chunks];
      const size_t nib_bytes = al16(n1) + al16(n_dq + 16) + al16(n_fq + 16);
      if (nib_bytes < 2 * al16(n1)) {
        h.nib = 1; h.n_dq = (int64_t)n_dq; h.n_fq = (int64_t)n_fq;
        h.off_nb = (uint32_t)o; uint8_t* nb = reinterpret_cast<uint8_t*>(base + o); o += al16(n1);
        h.off_dq = (uint32_t)o; uint8_t* dq = reinterpret_cast<uint8_t*>(base + o); o += al16(n_dq + 16);
        h.off_fq = (uint32_t)o; uint8_t* fq = reinterpret_cast<uint8_t*>(base + o); o += al16(n_fq + 16);
        par_for(n, n_threads, [&](int, int64_t ra, int64_t rb) {
          for (int64_t c = chunks_of(ra); c < chunks_of(rb); ++c) {
            uint32_t wd = cd[(size_t)c], wf = cf[(size_t)c];
            const int64_t e = std::min<int64_t>(n, (c + 1) * MCOV_BLOCK_CHUNK);
            for (int64_t i = c * MCOV_BLOCK_CHUNK; i < e; ++i) {
              uint8_t lo = dpos[i], hi = fc[i];
              if (lo >= 15) { dq[wd++] = lo; lo = 15; }
              if (hi >= 15) { fq[wf++] = hi; hi = 15; }
              nb[i] = (uint8_t)(lo | (hi << 4));
            }
          }
        });
      } else {
        h.nib = 0;
        h.off_dpos = (uint32_t)o; if (n > 0) std::memcpy(base + o, dpos, (size_t)n); o += al16(n1);
        h.off_fc = (uint32_t)o; if (n > 0) std::memcpy(base + o, fc, (size_t)n); o += al16(n1);
      }
    }
    // ---- escapes ----
    size_t n_esc = 0;
    for (auto& v : esc) n_esc += v.size();
    h.n_esc = (int64_t)n_esc;
    h.off_esc_idx = (uint32_t)o;
    uint32_t* qi = reinterpret_cast<uint32_t*>(base + o);
    o += al16(std::max<size_t>(n_esc, 1) * 4);
    h.off_esc_flag = (uint32_t)o;
    uint16_t* qf = reinterpret_cast<uint16_t*>(base + o);
    o += al16(std::max<size_t>(n_esc, 1) * 2);
    h.off_esc_cls = (uint32_t)o;
    uint8_t* qc = reinterpret_cast<uint8_t*>(base + o);
    o += al16(std::max<size_t>(n_esc, 1));
    { size_t k = 0; for (auto& v : esc) for (uint32_t i : v) { qi[k] = i; qf[k] = flag[i]; qc[k] = cls[i]; ++k; } }
    // ---- exceptions ----
    size_t n_exc = 0;
    for (auto& v : exc) n_exc += v.size();
    h.n_exc = (int64_t)n_exc;
    h.off_exc_idx = (uint32_t)o;
    uint32_t* ei = reinterpret_cast<uint32_t*>(base + o);
    o += al16(std::max<size_t>(n_exc, 1) * 4);
    h.off_exc_val = (uint32_t)o;
    int32_t* ev = reinterpret_cast<int32_t*>(base + o);
    o += al16(std::max<size_t>(n_exc, 1) * 4);
    { size_t k = 0; for (auto& v : exc) for (auto& e : v) { ei[k] = e.first; ev[k] = e.second; ++k; } }
    // ---- chunk table: everything a CTA of the unpack kernel needs to start at its chunk without looking at the others ----
    h.off_chunk = (uint32_t)o;
    {
      mcov_block_chunk* ct = reinterpret_cast<mcov_block_chunk*>(base + o);
      o += al16((size_t)std::max<int64_t>(n_chunks, 1) * sizeof(mcov_block_chunk));
      for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t c0 = c * MCOV_BLOCK_CHUNK;
        mcov_block_chunk& e = ct[c];
        e.dq_off = cd[(size_t)c]; e.fq_off = cf[(size_t)c]; e.op_off = cig_off[c0]; e.xop_off = cx[(size_t)c];
        e.pos_carry = c0 > 0 ? pos[c0 - 1] : 0;
        e.esc_first = 0xFFFFFFFFu; e.exc_first = 0xFFFFFFFFu; e.reserved = 0;
      }
      for (size_t k = 0; k < n_esc; ++k) if (k == 0 || qi[k - 1] / MCOV_BLOCK_CHUNK != qi[k] / MCOV_BLOCK_CHUNK) ct[qi[k] / MCOV_BLOCK_CHUNK].esc_first = (uint32_t)k;
      for (size_t k = 0; k < n_exc; ++k) if (k == 0 || ei[k - 1] / MCOV_BLOCK_CHUNK != ei[k] / MCOV_BLOCK_CHUNK) ct[ei[k] / MCOV_BLOCK_CHUNK].exc_first = (uint32_t)k;
    }
    // ---- mapq ----
    h.off_mapq = 0;
    if (mapq) { h.off_mapq = (uint32_t)o; if (n > 0) std::memcpy(base + o, mapq, (size_t)n); o += al16(n1); }
    if (o > (size_t)cap || o > 0xFFFFFFFFull) return MCOV_ERR_RANGE;     // (section offsets are 32-bit: a block holds < 4 GiB)
    h.total_bytes = (int64_t)o;
    std::memcpy(base, &h, sizeof(h));
    *bytes_out = (int64_t)o;
    return MCOV_OK;
  } catch (const std::bad_alloc&) { return MCOV_ERR_NOMEM; }
  catch (...) { return MCOV_ERR_ARG; }
}

}  // extern "C"
