// Host logic of the host-buffer pipeline (node-fhe-accelerate_b200/csrc/runtime.hpp: pipeline_chunk_sizes) checked on the
// CPU: every schedule covers the items exactly once with chunks no larger than the staging buffers, and the ramp is
// symmetric and only used when there is room for it.
#include <cstdio>
#include <numeric>

#include "../../node-fhe-accelerate_b200/csrc/runtime.hpp"

int main() {
    int bad = 0;
    const size_t item_counts[] = {0, 1, 2, 15, 16, 17, 127, 128, 129, 255, 256, 257, 300, 1000, 1024, 4096, 100000};
    const size_t chunks[] = {1, 2, 15, 16, 17, 64, 128, 1000, 1 << 20};
    for (size_t items : item_counts)
        for (size_t chunk : chunks)
            for (int ramp = 0; ramp < 2; ++ramp) {
                const std::vector<size_t> s = fheb::pipeline_chunk_sizes(items, chunk, ramp != 0);
                const size_t cap = chunk < items ? chunk : items;
                size_t sum = 0;
                bool ok = true;
                for (size_t c : s) {
                    ok = ok && c >= 1 && c <= cap;
                    sum += c;
                }
                ok = ok && sum == items && (items != 0 || s.empty());
                if (!ramp) {  // equal chunks, the last one possibly short
                    for (size_t i = 0; i + 1 < s.size(); ++i) ok = ok && s[i] == cap;
                } else if (s.size() >= 3 && s.front() < cap) {  // ramp present: mirrored, doubling, small first chunk
                    size_t h = 0;
                    while (h + 1 < s.size() && s[h] < s[h + 1] && s[h + 1] == 2 * s[h]) ++h;
                    ok = ok && s.front() == cap / 8 && s.back() == s.front();
                    for (size_t i = 0; i < h && i < s.size(); ++i) ok = ok && s[i] == s[s.size() - 1 - i];
                }
                if (!ok) {
                    std::printf("bad schedule: items %zu chunk %zu ramp %d ->", items, chunk, ramp);
                    for (size_t c : s) std::printf(" %zu", c);
                    std::printf("\n");
                    ++bad;
                }
            }
    // the bench's own case: 1024 polynomials of 128 KB, 16 MB chunks (the 32 left over by the ramp rides in front of the tail)
    const std::vector<size_t> s = fheb::pipeline_chunk_sizes(1024, 128, true);
    const std::vector<size_t> want = {16, 32, 64, 128, 128, 128, 128, 128, 128, 32, 64, 32, 16};
    if (s != want) {
        std::printf("unexpected schedule for 1024 x 128:");
        for (size_t c : s) std::printf(" %zu", c);
        std::printf("\n");
        ++bad;
    }
    if (bad) return 1;
    std::printf("PIPELINE SCHEDULE OK\n");
    return 0;
}
