// host_chain.h -- the serial half of the ciphertext commitment, for verifier-scale streams.
//
// AESAccumulatingHash (src/ciphertext_hasher.rs:1-34) is h <- AES_K(h ^ ct) over the whole stream:
// 2.98 G strictly dependent AES calls for the Groth16 verifier.  One dependent T-table AES costs a
// GPU ~0.46 us (10 rounds x shared-memory + shuffle latency), i.e. 23 minutes per stream however
// many SMs there are, while an AES-NI core folds the same stream in ~35 s.  GSV_CT_COMMIT_HOST keeps
// every gate hash on the GPU and drains the ciphertext ring over PCIe to host threads that fold the
// chains with AES-NI, concurrently with garbling -- the reference's own producer / hasher-thread
// split (src/circuit/modes/garble_mode.rs + examples/groth16_garble.rs).  The GPU-fused chain
// (GSV_CT_COMMIT) remains the mode for large batches, where thousands of chains run side by side.
#pragma once
#include <cstddef>
#include <cstdint>

namespace gsv {

// true when the host CPU has AES-NI (the mode refuses to start otherwise)
bool host_chain_available();

// h[i] <- AES_K(h[i] ^ block(p, i)) for p = 0..n_pos-1 in order, i = 0..n_inst-1 (up to 8 chains are
// interleaved to fill the AES pipeline).  block(p, i) = base + (p * pos_stride + i * inst_stride) * 16.
void host_chain_fold(uint8_t* h, const uint8_t* base, size_t pos_stride, size_t inst_stride, size_t n_pos,
                     uint32_t n_inst);


// The drain layout of GSV_CT_COMMIT_HOST: `n_quads` quads of chains, quad q's rows at base + q * quad_bytes,
// row p = the 16-byte blocks of the quad's four chains at stream position p (64 bytes).  h holds
// 4 * n_quads chain states.  One 512-bit load per quad and step with VAES; 128-bit AES-NI otherwise.
void host_chain_fold_quads(uint8_t* h, const uint8_t* base, size_t quad_bytes, size_t n_pos, uint32_t n_quads);


// n_inst independent streams given by pointer (gc_{i}.bin images in host memory, FileSource): positions
// [first, first + n_pos) of every stream are folded into h (n_inst states), up to four chains interleaved.
void host_chain_fold_streams(uint8_t* h, const uint8_t* const* streams, uint64_t first, size_t n_pos, uint32_t n_inst);

}  // namespace gsv
