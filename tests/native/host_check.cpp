// tests/native/host_check.cpp -- TEST HARNESS ONLY.
// Compiles the host/device-portable pieces of the product (huff_build.cuh, inflate_core.cuh) with g++ so
// their logic can be compared against the oracle on a machine without a GPU.  Nothing here ships:
// libb2f.so never contains or calls this file.
#include "../../libflate_b200/csrc/huff_build.cuh"
#include <cstring>
#include <vector>

using namespace b2f;

extern "C" {

void hc_code_lengths(const uint32_t *freq, int n, int cap, uint8_t *width) {
    static HuffWork W;
    huff_code_lengths(freq, n, cap, width, W);
}

// returns header bit count; litcode[288], distcode[32], hdr_words[kHdrWords]
uint32_t hc_block_codes(const uint32_t *hist, uint32_t *litcode, uint32_t *distcode, uint32_t *hdr_words) {
    static HuffWork W;
    return build_block_codes(hist, litcode, distcode, hdr_words, W);
}
void hc_fixed_codes(uint32_t *litcode, uint32_t *distcode) { build_fixed_codes(litcode, distcode); }

void hc_length_code(uint32_t len, uint32_t *out3) { length_code(len, out3[0], out3[1], out3[2]); }
void hc_dist_code(uint32_t dist, uint32_t *out3) { dist_code(dist, out3[0], out3[1], out3[2]); }
uint32_t hc_hdr_words(void) { return kHdrWords; }

}  // extern "C"
