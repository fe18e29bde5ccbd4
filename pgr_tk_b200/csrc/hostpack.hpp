// hostpack.hpp — host side of the packed upload (ctx.cu: upload_packed): bases leave the host as three bit planes per
// 32-byte block of the device sequence store (3 bits per base instead of 8 over PCIe) and are expanded to canonical
// ASCII by unpack_kernel on the device.  Only the letter classes the reference's LUT distinguishes survive
// (shmmrutils.rs:426-436: A/a/0, C/c/1, G/g/2, T/t/3, everything else), which is all sequence_to_shmmrs reads.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace pgr {

// Planes of block b (32 consecutive bytes of the store), bit j = byte j of the block:
//   v = 1 : a base, code = p1:p0 (A 0, C 1, G 2, T 3); the padding behind the end of a sequence in its last block counts as
//           code 0 (those bytes are never read as sequence)
//   v = 0 : a byte the reference maps to no base (written back as 'N'); p0 = p1 = 0
// Packs n_bytes bytes starting at src (the beginning of a block); returns the AND of the validity words written (all ones:
// every byte was a base, and the validity plane need not cross PCIe).
uint32_t pack_bases(const uint8_t *src, size_t n_bytes, uint32_t *p0, uint32_t *p1, uint32_t *v);
const char *pack_isa();   // "avx512bw", "avx2" or "scalar": what pack_bases runs on this machine

// persistent worker pool shared by the library's host-side loops; fn(i) for i in [0, n), the caller takes part
void parallel_for(size_t n, const std::function<void(size_t)> &fn);
unsigned pool_threads();  // workers + caller: PGR_B200_HOST_THREADS, else the CPUs this process may run on divided by LOCAL_WORLD_SIZE, at most 32

}  // namespace pgr
