// Lane-serial LZ4 block decoder: ONE THREAD decodes one independent block.
//
// An LZ4 block is a serial chain (every sequence starts where the previous one ended), so a warp that cooperates
// on one block executes ~150 warp instructions per sequence with most lanes redundant (profiles/: the warp-per-block
// decoder is issue-bound at ~2 warp instructions per output byte). Here 32 lanes of a warp walk 32 different blocks;
// the loop body is a small state machine whose steps are the same code for every lane (token, match header, one
// 8-byte copy / 16-byte fill, flush), so the lanes stay converged while sitting in different sequences.
//
// Per lane, 32 words of shared memory, interleaved over the CTA (word j of thread t at smem[j*T + t]: every access
// of a lane hits bank t%32, so arbitrary per-lane indices are conflict free):
//   words [0,16)  : ring with 64 bytes of the compressed stream (4 chunks of 16 bytes, the next chunk prefetched in
//                   registers)
//   words [16,32) : ring with the last 64 output bytes. Matches with offset <= kLaneNear read it; complete 16-byte
//                   chunks are written to global memory with one 128-bit store; farther matches read the block's own
//                   flushed output back through L2.
// Copies move 8 bytes per step through registers (3 aligned ring words in, funnel shift, 3 aligned ring words out);
// overlapping matches double their offset (after copying `off` bytes the region is periodic with 2*off); runs with
// period 1, 2 or 4 are filled 16 bytes per step.
//
// This file is compiled twice: by nvcc into lz4_decode_lanes_kernel (lz4_decode.cu) and by g++ into
// tools/lane_sim.cpp, which checks the state machine against liblz4-compressed blocks without a GPU.
// The includer defines SQYB_LANE_FN, kLaneStride and the lane_* primitives listed below.
//
//   uint32_t lane_funnel_r(lo, hi, sh), lane_funnel_l(lo, hi, sh)       32-bit funnel shifts, sh in [0,32)
//   int      lane_ffs(x)                                                 1-based index of the lowest set bit
//   uint4    lane_load_chunk(const uint8_t* addr, sbeg, send)           16 compressed bytes, zero outside [sbeg,send)
//   uint32_t lane_load_out32(const uint8_t* p)                          4-byte aligned read of flushed output
//   void     lane_store_out16(uint8_t* p, uint4 v)                      16-byte aligned store
#pragma once

constexpr uint32_t kLaneNear = 40;   // offsets up to this are served by the output ring

enum LaneMode : uint32_t { kLaneTok = 0, kLaneLit = 2, kLaneNearCopy = 3, kLaneFarCopy = 4, kLaneIdle = 5 };

enum LaneErr : uint32_t { kLaneOk = 0, kLaneBadBlock = 4, kLaneSizeMismatch = 5 };

struct Lane {
  const uint8_t* A;     // 16-byte aligned address at or below the block's first compressed byte
  uint8_t* d;           // block output (16-byte aligned)
  uint32_t ip, end;     // positions in the compressed stream, relative to A
  uint32_t op, dcap;    // output position, decoded size promised by the block table
  uint32_t rem, off;    // bytes left in the current copy, match offset (doubles on overlapping copies)
  uint32_t cur;         // first chunk resident in the compressed ring
  uint32_t acc;         // bytes below op of the output word that holds op (zero above)
  uint32_t mode;        // LaneMode | run flag << 4 | token match nibble << 8
  uint4 pre;            // chunk cur+4
};

SQYB_LANE_FN uint32_t lane_cbyte(const uint32_t* base, uint32_t p) {
  return (base[((p >> 2) & 15u) * kLaneStride] >> ((p & 3u) * 8u)) & 0xffu;
}

// 8 bytes starting at byte position p of the ring whose first word is R (0 compressed, 16 output)
SQYB_LANE_FN void lane_read8(const uint32_t* base, uint32_t R, uint32_t p, uint32_t& lo, uint32_t& hi) {
  const uint32_t wi = p >> 2, sh = (p & 3u) * 8u;
  const uint32_t w0 = base[(R + (wi & 15u)) * kLaneStride];
  const uint32_t w1 = base[(R + ((wi + 1u) & 15u)) * kLaneStride];
  const uint32_t w2 = base[(R + ((wi + 2u) & 15u)) * kLaneStride];
  lo = lane_funnel_r(w0, w1, sh);
  hi = lane_funnel_r(w1, w2, sh);
}

// appends k (1..8) bytes of lo:hi at output position op. Bytes above k are don't-care: they land on positions that
// are written again before they are read and at most 11 bytes ahead, i.e. on ring slots older than op-52.
SQYB_LANE_FN void lane_write8(uint32_t* base, uint32_t op, uint32_t lo, uint32_t hi, uint32_t k, uint32_t& acc) {
  const uint32_t wi = op >> 2, sh = (op & 3u) * 8u;
  const uint32_t W0 = acc | (lo << sh);
  const uint32_t W1 = lane_funnel_l(lo, hi, sh);
  const uint32_t W2 = lane_funnel_l(hi, 0u, sh);
  base[(16u + (wi & 15u)) * kLaneStride] = W0;
  base[(16u + ((wi + 1u) & 15u)) * kLaneStride] = W1;
  base[(16u + ((wi + 2u) & 15u)) * kLaneStride] = W2;
  const uint32_t np = op + k, dw = (np >> 2) - wi;
  const uint32_t ws = dw == 0u ? W0 : (dw == 1u ? W1 : W2);
  acc = ws & ((1u << ((np & 3u) * 8u)) - 1u);
}

// chunk cur leaves the compressed ring, the prefetched chunk cur+4 takes its slot, chunk cur+5 is requested
SQYB_LANE_FN void lane_advance(Lane& L, uint32_t* base, const uint8_t* sbeg, const uint8_t* send) {
  const uint32_t s = (L.cur & 3u) * 4u;
  base[(s + 0u) * kLaneStride] = L.pre.x;
  base[(s + 1u) * kLaneStride] = L.pre.y;
  base[(s + 2u) * kLaneStride] = L.pre.z;
  base[(s + 3u) * kLaneStride] = L.pre.w;
  L.cur++;
  L.pre = lane_load_chunk(L.A + 16ull * (L.cur + 4u), sbeg, send);
}

// starts a block: s = first compressed byte, csize bytes; d = output (16-byte aligned), dsize bytes expected
SQYB_LANE_FN void lane_begin(Lane& L, uint32_t* base, const uint8_t* s, uint32_t csize, uint8_t* d, uint32_t dsize,
                             const uint8_t* sbeg, const uint8_t* send) {
  L.A = s - ((uintptr_t)s & 15u);
  L.d = d;
  L.ip = (uint32_t)(s - L.A);
  L.end = L.ip + csize;
  L.op = 0;
  L.dcap = dsize;
  L.rem = 0;
  L.off = 0;
  L.cur = 0;
  L.acc = 0;
  L.mode = kLaneTok;
#pragma unroll
  for (uint32_t c = 0; c < 4; ++c) {
    const uint4 v = lane_load_chunk(L.A + 16u * c, sbeg, send);
    base[(4u * c + 0u) * kLaneStride] = v.x;
    base[(4u * c + 1u) * kLaneStride] = v.y;
    base[(4u * c + 2u) * kLaneStride] = v.z;
    base[(4u * c + 3u) * kLaneStride] = v.w;
  }
  L.pre = lane_load_chunk(L.A + 64u, sbeg, send);
  if (csize == 0) L.mode = kLaneIdle;   // caller checks dsize == 0
}

// length bytes (255, 255, ..., <255) at ip. Up to three 255s and the terminator are taken from one 4-byte read; longer
// chains (lengths >= 780) walk byte by byte and advance the ring. Returns false when the stream ends inside the chain.
SQYB_LANE_FN bool lane_length_bytes(Lane& L, uint32_t* base, uint32_t& v, const uint8_t* sbeg, const uint8_t* send) {
  uint32_t lo, hi;
  lane_read8(base, 0u, L.ip, lo, hi);
  const uint32_t x = ~lo;                       // a non-zero byte of x marks a length byte below 255
  bool ok;
  if (x != 0u) {
    const uint32_t n = (uint32_t)(lane_ffs(x) - 1) >> 3;   // 255s in front of the terminator
    v += 255u * n + ((lo >> (8u * n)) & 0xffu);
    L.ip += n + 1u;
    ok = L.ip <= L.end;
  } else {
    uint32_t b = 255u;
    while (b == 255u && L.ip < L.end) {
      b = lane_cbyte(base, L.ip);
      L.ip++;
      v += b;
      if ((L.ip >> 4) != L.cur) lane_advance(L, base, sbeg, send);
    }
    ok = b != 255u;
  }
  return ok;
}

SQYB_LANE_FN void lane_flush_chunk(Lane& L, const uint32_t* base, uint32_t c) {
  const uint32_t c4 = c * 4u;
  uint4 v;
  v.x = base[(16u + (c4 & 15u)) * kLaneStride];
  v.y = base[(16u + ((c4 + 1u) & 15u)) * kLaneStride];
  v.z = base[(16u + ((c4 + 2u) & 15u)) * kLaneStride];
  v.w = base[(16u + ((c4 + 3u) & 15u)) * kLaneStride];
  lane_store_out16(L.d + (size_t)c * 16u, v);
}

// One iteration of the lane's block: [token] -> [up to 8 literal bytes] -> [match header] -> [8 match bytes or a 16-byte
// fill] -> flush. A short sequence completes in one iteration; long literal runs and long matches continue in the
// following iterations (rem > 0) and skip the front part. Returns kLaneOk or an error; L.mode == kLaneIdle afterwards
// means the block is finished (or failed). An iteration consumes at most 15 compressed bytes (+ long length chains,
// which advance the ring themselves), so one ring advance at the end keeps 32+ bytes of look-ahead resident.
// Single exit, no early returns: the lanes of a warp sit in different sequences, and every branch has to reconverge
// right behind its `if` for the warp to share the instructions of the common path.
SQYB_LANE_FN uint32_t lane_step(Lane& L, uint32_t* base, const uint8_t* sbeg, const uint8_t* send) {
  uint32_t m = L.mode & 15u;
  bool ok = true;
  const uint32_t op0 = L.op;
  if (m == kLaneTok) {
    // (a block ends with literals, never with a match: ip == end here is malformed)
    const uint32_t token = lane_cbyte(base, L.ip);
    ok = L.ip < L.end;
    L.ip++;
    uint32_t lit = token >> 4;
    if (lit == 15u) ok = lane_length_bytes(L, base, lit, sbeg, send) && ok;
    ok = ok && L.ip + lit <= L.end && L.op + lit <= L.dcap;
    L.rem = lit;
    m = kLaneLit;
    L.mode = m | ((token & 15u) << 8);
  }
  if (m == kLaneLit && ok) {
    if (L.rem) {
      const uint32_t k = L.rem < 8u ? L.rem : 8u;
      uint32_t lo, hi;
      lane_read8(base, 0u, L.ip, lo, hi);
      lane_write8(base, L.op, lo, hi, k, L.acc);
      L.op += k;
      L.ip += k;
      L.rem -= k;
    }
    if (L.rem == 0u) {
      if (L.ip >= L.end) {
        m = kLaneIdle;                                    // the block ends with literals
      } else {
        uint32_t lo, hi;
        lane_read8(base, 0u, L.ip, lo, hi);
        const uint32_t off = lo & 0xffffu;
        ok = L.ip + 2u <= L.end;
        L.ip += 2u;
        uint32_t mlen = (L.mode >> 8) & 15u;
        if (mlen == 15u) ok = lane_length_bytes(L, base, mlen, sbeg, send) && ok;
        mlen += 4u;
        ok = ok && off != 0u && off <= L.op && L.op + mlen <= L.dcap;
        L.off = off;
        L.rem = mlen;
        m = off <= kLaneNear ? kLaneNearCopy : kLaneFarCopy;
        L.mode = m | ((off == 1u || off == 2u || off == 4u) ? 16u : 0u);
      }
    }
  }
  if ((m == kLaneNearCopy || m == kLaneFarCopy) && ok) {
    const bool run = (L.mode & 16u) != 0u;
    if (m == kLaneNearCopy && run && (L.op & 3u) == 0u && L.rem >= 16u && L.off >= 4u) {
      // period divides 4 and (off >= 4) the four bytes in front of op already repeat it: every aligned word from here
      // on equals the aligned word in front of op
      const uint32_t wi = L.op >> 2;
      const uint32_t w = base[(16u + ((wi - 1u) & 15u)) * kLaneStride];
      base[(16u + (wi & 15u)) * kLaneStride] = w;
      base[(16u + ((wi + 1u) & 15u)) * kLaneStride] = w;
      base[(16u + ((wi + 2u) & 15u)) * kLaneStride] = w;
      base[(16u + ((wi + 3u) & 15u)) * kLaneStride] = w;
      L.op += 16u;
      L.rem -= 16u;
    } else {
      uint32_t k = L.rem < 8u ? L.rem : 8u;
      uint32_t lo, hi;
      if (m == kLaneFarCopy) {
        const uint32_t sp = L.op - L.off, a = sp & ~3u, sh = (sp & 3u) * 8u;
        const uint32_t w0 = lane_load_out32(L.d + a), w1 = lane_load_out32(L.d + a + 4u), w2 = lane_load_out32(L.d + a + 8u);
        lo = lane_funnel_r(w0, w1, sh);
        hi = lane_funnel_r(w1, w2, sh);
      } else {
        if (k > L.off) k = L.off;                                     // never read a byte written in this step
        if (run && (L.op & 3u) && k > 4u - (L.op & 3u)) k = 4u - (L.op & 3u);   // reach word alignment, then fill
        lane_read8(base, 16u, L.op - L.off, lo, hi);
      }
      lane_write8(base, L.op, lo, hi, k, L.acc);
      L.op += k;
      L.rem -= k;
      if (k == L.off) L.off <<= 1;                                    // [op-2*off, op) is periodic now
    }
    if (L.rem == 0u) L.mode = kLaneTok;
  }
  uint32_t rc = ok ? kLaneOk : kLaneBadBlock;
  if (ok) {
    // at most 8 literal + 16 match bytes were written: up to two 16-byte chunks became complete
    for (uint32_t c = op0 >> 4; c < (L.op >> 4); ++c) lane_flush_chunk(L, base, c);
    if (m == kLaneIdle) {
      // write the incomplete last chunk and check the promised size
      for (uint32_t p = L.op & ~15u; p < L.op; ++p)
        L.d[p] = (uint8_t)(base[(16u + ((p >> 2) & 15u)) * kLaneStride] >> ((p & 3u) * 8u));
      if (L.op != L.dcap || L.ip != L.end) rc = kLaneSizeMismatch;
    }
  }
  if (!ok || m == kLaneIdle) L.mode = kLaneIdle;
  if ((L.ip >> 4) != L.cur) lane_advance(L, base, sbeg, send);
  return rc;
}
