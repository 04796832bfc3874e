// inflate.cuh -- inflate side of the path (inflate.zig, huffman_decoder.zig, bit_reader.zig,
// CircularBuffer.zig) and container checksums (container.zig:168-206) on the device.
#pragma once
#include "common.cuh"

namespace fb {

struct MemberDesc {
    uint64_t in_off, in_len;    // compressed member inside d_in
    uint64_t out_off, out_cap;  // where its plain bytes go inside d_out
    uint64_t hist;              // bytes of earlier output directly before out_off the member may reference
    uint32_t start_bit;         // decoding starts at this bit of the first byte (a member resumed at a block boundary)
    uint32_t flags;             // kMember*
};
// A streaming decompressor decodes a member piece by piece (inflate_par.cu only):
//   kMemberResume    the container header was consumed by an earlier call: start with a block header
//   kMemberPartial   more input may follow what is given: whatever stops the decode inside a block (input or output
//                    running out, or an error that zero-padded look-ahead may have faked) ends the call at the start
//                    of that block with status OK, so that it can be resumed there with more input
//   kMemberNoFooter  leave the container footer to the caller and return the checksum of the bytes produced
enum : uint32_t { kMemberResume = 1, kMemberPartial = 2, kMemberNoFooter = 4 };
struct MemberResult {
    uint64_t out_len, consumed;
    uint32_t status, pad;
    uint64_t resume_bits;       // bit offset from in_off of the block boundary the call stopped at
    uint32_t sum;               // kMemberNoFooter: CRC-32 (gzip) / Adler-32 (zlib) of the out_len bytes produced
    uint32_t info;              // bit 0: the final block was decoded; bits 8..15: the status that ended a partial call early
};
constexpr uint32_t kNeedsHistory = 100;  // internal: match reaches before the member; redo after predecessors

// one CTA per member, lane-parallel symbol decode inside the member (inflate_par.cu): the default
// d_scratch: inflate_par_scratch_bytes(k, sm_count) bytes of device memory (work counter + one match queue per CTA)
size_t inflate_par_scratch_bytes(uint32_t k, int sm_count);
cudaError_t inflate_members_par(int container, const uint8_t* d_in, const MemberDesc* d_desc, uint32_t k, uint8_t* d_out,
                                MemberResult* d_res, void* d_scratch, int sm_count, cudaStream_t st, bool pieces = false);
// one warp per member (inflate.cu): kept for comparison, FB200_INFLATE=warp
cudaError_t inflate_members(int container, const uint8_t* d_in, const MemberDesc* d_desc, uint32_t k, uint8_t* d_out,
                            MemberResult* d_res, cudaStream_t st);

// Positions of d_data[0..n) that look like the start of a gzip member (ID1 ID2 CM = 1f 8b 08, reserved flag bits zero,
// container.zig:119-126): up to `cap` of them, unordered, into d_list; *d_count receives how many there are.
cudaError_t gzip_candidates_device(const uint8_t* d_data, uint64_t n, uint64_t* d_list, uint32_t cap, uint32_t* d_count, cudaStream_t st);

// CRC-32 (IEEE, reflected) / Adler-32 of d_data[0..n) into *d_result (device u32)
cudaError_t crc32_device(const uint8_t* d_data, uint64_t n, uint32_t* d_result, cudaStream_t st);
cudaError_t adler32_device(const uint8_t* d_data, uint64_t n, uint32_t* d_result, uint64_t* d_scratch2, cudaStream_t st);

}  // namespace fb
