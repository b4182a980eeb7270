// CPU test of bioinfo1_b200/csrc/host_pack.hpp (the 2-bit packer of the pointer-array entry point): the SSE2 routine against a
// byte-at-a-time restatement of pack_kernel's layout and flag rules, on random sequences with and without foreign bytes.
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../bioinfo1_b200/csrc/host_pack.hpp"

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main() {
    const std::string alphabets[] = {"ACGT", "ACGTN", "ACGT-", "ACGTacgt", std::string("ACGT\0\xff", 6), "A", "T"};
    long checked = 0;
    for (int it = 0; it < 200000; ++it) {
        const std::string& ab = alphabets[rnd() % 7];
        const uint32_t len = (uint32_t)(rnd() % (it % 50 == 0 ? 4000 : 200));
        std::string s(len, 'A');
        for (auto& ch : s) ch = ab[rnd() % ab.size()];
        if (it % 3 == 0 && len) { const uint32_t at = (uint32_t)(rnd() % len); s[at] = "N-nE\0U"[rnd() % 6]; }   // one foreign byte anywhere
        std::vector<uint32_t> got(len / 16 + 2, 0xdeadbeefu), exp(len / 16 + 2, 0u);
        const uint8_t flag = b200::host_pack_sequence(s.data(), len, got.data());
        uint8_t eflag = 0;
        for (uint32_t i = 0; i < len; ++i) {
            const unsigned char c = (unsigned char)s[i];
            exp[i / 16] |= (uint32_t)((c >> 1) & 3u) << (2 * (i % 16));
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') eflag |= 2;
            if (c == '-') eflag |= 1;
        }
        if (eflag == 0 && got != exp) { std::printf("FAIL words len %u\n", len); return 1; }   // (flagged sequences are not read as 2-bit words)
        if (flag != eflag) { std::printf("FAIL flag len %u got %u want %u\n", len, flag, eflag); return 1; }
        ++checked;
    }
    std::printf("host_pack ok: %ld sequences\n", checked);
    return 0;
}
