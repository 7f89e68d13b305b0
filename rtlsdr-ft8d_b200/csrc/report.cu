// report.cu -- the step AFTER the hot path: turning decoder_results records into the
// daemon's three output formats.  Host code only (a few hundred bytes per slot; there is nothing here for a GPU to do):
//   * PSKreporter IPFIX datagram           replaces postSpots(),        /root/reference/rtlsdr_ft8d.c:365-582
//   * web-cluster POST form fields         replaces webClusterSpots(),  /root/reference/rtlsdr_ft8d.c:586-632
//   * the console table                    replaces printSpots(),       /root/reference/rtlsdr_ft8d.c:635-663
// No sockets, no curl: the functions produce the bytes / strings; sending them stays with the caller (the reference's
// own postSpots() is disabled upstream by a `return` on its first line).  Byte-for-byte parity with the reference's
// packet-building code is pinned in tests/test_report.py against oracle/_ref/libref_report.so.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>

#include "../../include/ft8b200.h"

namespace {

// Big-endian writer over a caller-supplied buffer; keeps counting past the end so the needed size is always known.
struct Writer {
    uint8_t *p;
    size_t cap, n = 0;
    Writer(uint8_t *out, size_t cap_) : p(out), cap(cap_) {}
    void u8(uint32_t v) { if (n < cap) p[n] = (uint8_t)v; ++n; }
    void u16(uint32_t v) { u8(v >> 8); u8(v); }
    void u32(uint32_t v) { u16(v >> 16); u16(v); }
    void bytes(const char *s, size_t len) { for (size_t k = 0; k < len; ++k) u8((uint8_t)s[k]); }
    void lstr(const char *s) { const size_t len = strlen(s); u8((uint32_t)len); bytes(s, len); }  // 1-byte length + text
    void pad4() { while (n & 3) u8(0); }
    void patch16(size_t at, uint32_t v) { if (at + 1 < cap) { p[at] = (uint8_t)(v >> 8); p[at + 1] = (uint8_t)v; } }
    bool ok() const { return n <= cap; }
};

// PSKreporter information elements (enterprise 30351 = 0x768F).  len 0xFFFF = variable length.
struct Field { uint16_t id, len; bool enterprise; };
constexpr uint32_t kEnterprise = 0x0000768Fu;
constexpr Field kReceiverFields[] = {{0x8002, 0xFFFF, true},   // receiverCallsign
                                     {0x8004, 0xFFFF, true},   // receiverLocator
                                     {0x8008, 0xFFFF, true}};  // decodingSoftware
constexpr Field kSenderFields[] = {{0x8001, 0xFFFF, true},     // senderCallsign
                                   {0x8005, 4, true},          // frequency
                                   {0x8006, 1, true},          // sNR
                                   {0x800A, 0xFFFF, true},     // mode
                                   {0x8003, 0xFFFF, true},     // senderLocator
                                   {0x800B, 1, true},          // informationSource
                                   {0x0096, 4, false}};        // flowStartSeconds
constexpr uint16_t kReceiverLink = 0x9992, kSenderLink = 0x9993;

template <int N>
void put_template(Writer &w, uint16_t set_id, uint16_t link, const Field (&f)[N], bool scope_count) {
    const size_t start = w.n;
    w.u16(set_id);
    w.u16(0);  // set length, patched below
    w.u16(link);
    w.u16(N);
    if (scope_count) w.u16(0);  // options template: scope field count
    for (const Field &e : f) {
        w.u16(e.id);
        w.u16(e.len);
        if (e.enterprise) w.u32(kEnterprise);
    }
    w.pad4();
    w.patch16(start + 2, (uint32_t)(w.n - start));
}

size_t bounded_len(const char *s, size_t cap) { size_t n = 0; while (n < cap && s[n]) ++n; return n; }

// decoder_results strings are fixed char arrays; a record copied from the device is always terminated, but a caller's
// might not be: read at most the array.
void put_field(Writer &w, const char *s, size_t cap) {
    const size_t len = bounded_len(s, cap);
    w.u8((uint32_t)len);
    w.bytes(s, len);
}

}  // namespace

extern "C" {

const char *ft8b200_report_app_version(void) { return "ft8b200_v0.1"; }

int ft8b200_pskreporter_datagram(const struct decoder_results *spots, uint32_t n_spots, const struct decoder_options *station, const char *app_version,
                                 uint32_t unixtime, uint32_t sequence, uint32_t random_id, uint8_t *out, size_t cap, uint32_t *n_reported) {
    if (n_reported) *n_reported = 0;
    if ((!spots && n_spots) || !station || !out) return -1;
    if (!app_version) app_version = ft8b200_report_app_version();
    if (strlen(app_version) > 200) return -1;  // the reference's receiver block is a 256-byte array (rtlsdr_ft8d.c:453)
    Writer w(out, cap);
    // message header (rtlsdr_ft8d.c:438-450): version 10, total length (patched last), export time, sequence, observation domain
    w.u16(0x000A);
    w.u16(0);
    w.u32(unixtime);
    w.u32(sequence);
    w.u32(random_id);
    put_template(w, 3, kReceiverLink, kReceiverFields, true);   // :381-396
    put_template(w, 2, kSenderLink, kSenderFields, false);      // :398-423
    // receiver record (:453-483)
    size_t start = w.n;
    w.u16(kReceiverLink);
    w.u16(0);
    put_field(w, station->rcall, sizeof(station->rcall));
    put_field(w, station->rloc, sizeof(station->rloc));
    w.lstr(app_version);
    w.pad4();
    w.patch16(start + 2, (uint32_t)(w.n - start));
    // sender records (:486-538): the reference stops adding spots once the set has grown past 1200 bytes
    start = w.n;
    w.u16(kSenderLink);
    w.u16(0);
    uint32_t used = 0;
    for (; used < n_spots && w.n - start <= 1200; ++used) {
        const decoder_results &r = spots[used];
        put_field(w, r.call, sizeof(r.call));
        w.u32((uint32_t)r.freq + station->freq);
        w.u8((uint32_t)(int32_t)((int8_t)r.snr - 20));  // (int8_t)snr - 20, truncated to one byte (:511)
        w.lstr("FT8");
        put_field(w, r.loc, sizeof(r.loc));
        w.u8(1);  // information source: automatically extracted
        w.u32(unixtime);
    }
    w.pad4();
    w.patch16(start + 2, (uint32_t)(w.n - start));
    w.patch16(2, (uint32_t)w.n);
    if (!w.ok()) return -1;
    if (n_reported) *n_reported = used;
    return (int)w.n;
}

int ft8b200_pskreporter_batch(const struct decoder_results *spots, const int32_t *n_spots, int n_slots, int max_messages,
                              const struct decoder_options *station, const char *app_version, const uint32_t *unixtime, uint32_t first_sequence,
                              uint32_t random_id, uint8_t *out, size_t stride, int32_t *lengths) {
    if (!spots || !n_spots || !out || !lengths || !unixtime || n_slots < 0 || max_messages < 0) return -1;
    int sent = 0;
    for (int s = 0; s < n_slots; ++s) {
        lengths[s] = 0;
        int32_t n = n_spots[s];
        if (n <= 0) continue;  // the daemon has nothing to report for an empty slot
        if (n > max_messages) n = max_messages;
        const int len = ft8b200_pskreporter_datagram(spots + (size_t)s * max_messages, (uint32_t)n, station, app_version, unixtime[s],
                                                     first_sequence + (uint32_t)sent, random_id, out + (size_t)s * stride, stride, nullptr);
        if (len < 0) return -1;
        lengths[s] = len;
        ++sent;
    }
    return sent;
}

int ft8b200_webcluster_form(const struct decoder_results *spot, const struct decoder_options *station, ft8b200_cluster_form_t *form) {
    if (!spot || !station || !form) return -1;
    char call[sizeof(spot->call) + 1] = {0}, loc[sizeof(spot->loc) + 1] = {0}, rcall[sizeof(station->rcall) + 1] = {0}, rloc[sizeof(station->rloc) + 1] = {0};
    memcpy(call, spot->call, bounded_len(spot->call, sizeof(spot->call)));
    memcpy(loc, spot->loc, bounded_len(spot->loc, sizeof(spot->loc)));
    memcpy(rcall, station->rcall, bounded_len(station->rcall, sizeof(station->rcall)));
    memcpy(rloc, station->rloc, bounded_len(station->rloc, sizeof(station->rloc)));
    memset(form, 0, sizeof(*form));
    snprintf(form->mycall, sizeof(form->mycall), "%s", rcall);
    // kHz with float arithmetic, as the reference computes it: 14 074 800 Hz prints as "14074.799" (:604)
    snprintf(form->freq, sizeof(form->freq), "%8f", (double)((float)((uint32_t)spot->freq + station->freq) / 1000.0f));
    snprintf(form->dxcall, sizeof(form->dxcall), "%s", call);
    snprintf(form->info, sizeof(form->info), "M2M FT8 [%s - %s]", rloc, loc);
    return 0;
}

int ft8b200_format_spots(const struct decoder_results *spots, uint32_t n_spots, uint32_t dial_freq, uint32_t unixtime, char *out, size_t cap) {
    if ((!spots && n_spots) || (!out && cap)) return -1;
    size_t n = 0;
    auto add = [&](const char *fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        const int k = vsnprintf(n < cap ? out + n : nullptr, n < cap ? cap - n : 0, fmt, ap);
        va_end(ap);
        if (k > 0) n += (size_t)k;
    };
    if (n_spots == 0) {
        const time_t t = (time_t)unixtime;
        struct tm g;
        gmtime_r(&t, &g);
        add("No spot %04d-%02d-%02d %02d:%02dz\n", g.tm_year + 1900, g.tm_mon + 1, g.tm_mday, g.tm_hour, g.tm_min);
        return (int)n;
    }
    add("  Score     Freq       Call    Loc\n");
    for (uint32_t k = 0; k < n_spots; ++k) {
        char call[sizeof(spots[k].call) + 1] = {0}, loc[sizeof(spots[k].loc) + 1] = {0};
        memcpy(call, spots[k].call, bounded_len(spots[k].call, sizeof(spots[k].call)));
        memcpy(loc, spots[k].loc, bounded_len(spots[k].loc, sizeof(spots[k].loc)));
        // the reference passes (int32 + uint32) = an unsigned value to %8d: it prints as the same bits, signed
        add("     %2d %8d %10s %6s\n", spots[k].snr, (int32_t)((uint32_t)spots[k].freq + dial_freq), call, loc);
    }
    return (int)n;  // characters needed (excluding the terminator), like snprintf
}

}  // extern "C"
