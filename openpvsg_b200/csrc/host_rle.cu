// Host-side half of the tube wire format (runs on the CPU; plain C++ in the same library so the
// Python layer stays a thin binding): turns the run-boundary events produced by pvsg_rle_events
// into pycocotools-compatible RLE strings, one per kept segment -- stable counting sort of the
// events by segment slot, run lengths = position differences, then `rleToString` (5 payload bits
// per character, continuation bit, deltas against the count two places back from the 4th count on;
// reference: pycocotools maskApi.c rleToString, used through models/unitrack/utils/io.py:14-37).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pvsg.h"

namespace {
inline int64_t put_count(char* out, int64_t pos, int64_t cap, long x) {
    bool more = true;
    while (more) {
        long c = x & 0x1f;
        x >>= 5;
        more = (c & 0x10) ? x != -1 : x != 0;
        if (more) c |= 0x20;
        if (pos >= cap) return -1;
        out[pos++] = (char)(c + 48);
    }
    return pos;
}
}  // namespace

extern "C" int64_t pvsg_rle_strings_host(const uint32_t* ev_pos, const int16_t* ev_slot, int64_t n, int nseg,
                                         uint32_t hw, char* out, int64_t out_cap, int64_t* seg_off) {
    if (!ev_pos || !ev_slot || !out || !seg_off || n < 0 || nseg <= 0) return PVSG_ERR_INVALID_ARG;
    int64_t* start = (int64_t*)calloc((size_t)nseg + 1, sizeof(int64_t));
    uint32_t* sorted = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    if (!start || !sorted) { free(start); free(sorted); return PVSG_ERR_LAUNCH; }
    for (int64_t i = 0; i < n; ++i) {
        const int s = ev_slot[i];
        if (s < 0 || s >= nseg) { free(start); free(sorted); return PVSG_ERR_INVALID_ARG; }
        ++start[s + 1];
    }
    for (int k = 0; k < nseg; ++k) start[k + 1] += start[k];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)nseg);
    if (!fill) { free(start); free(sorted); return PVSG_ERR_LAUNCH; }
    memcpy(fill, start, sizeof(int64_t) * (size_t)nseg);
    for (int64_t i = 0; i < n; ++i) sorted[fill[ev_slot[i]]++] = ev_pos[i];   // stable: walk order kept
    int64_t pos = 0;
    for (int k = 0; k < nseg && pos >= 0; ++k) {
        seg_off[k] = pos;
        long c1 = 0, c2 = 0;      // counts one and two places back
        int64_t m = 0;            // index of the count being written
        uint32_t prev = 0;
        const int64_t b = start[k], e = start[k + 1];
        for (int64_t i = b; i <= e && pos >= 0; ++i) {
            const uint32_t p = i < e ? sorted[i] : hw;
            const long c = (long)p - (long)prev;
            prev = p;
            pos = put_count(out, pos, out_cap, m > 2 ? c - c2 : c);
            c2 = c1; c1 = c;
            ++m;
        }
    }
    free(start); free(sorted); free(fill);
    if (pos < 0) return PVSG_ERR_UNSUPPORTED;   // out_cap too small
    seg_off[nseg] = pos;
    return pos;
}
