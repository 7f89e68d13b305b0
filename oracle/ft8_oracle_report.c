/* oracle/ft8_oracle_report.c -- CPU restatement of the daemon's reporting formats.
 *
 * TEST INFRASTRUCTURE ONLY (see ft8_oracle.h): the product never includes, links or calls this file.
 * Pinned against the reference's own packet-building code (oracle/_ref/libref_report.so, see ref_report_harness.c)
 * by tests/test_report.py, and against tests/golden/report.npz (generated from it by tools/make_golden.py).
 *
 * Plain scalar C in the reference's statement order: every block is laid out in its own scratch array, sizes are
 * fixed up afterwards, and the pieces are concatenated -- as postSpots() does.
 */
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "ft8_oracle.h"

#pragma GCC diagnostic ignored "-Wformat-truncation" /* the reference truncates into its fixed buffers on purpose (rtlsdr_ft8d.c:599-606) */

static void be16(unsigned char *p, uint32_t v) { p[0] = (unsigned char)(v >> 8); p[1] = (unsigned char)v; }
static void be32(unsigned char *p, uint32_t v) { be16(p, v >> 16); be16(p + 2, v); }

/* length-prefixed string, the idiom of rtlsdr_ft8d.c:463-466 */
static uint32_t pstr(unsigned char *dst, uint32_t at, const char *s) {
    size_t n = strlen(s);
    dst[at] = (unsigned char)n;
    memcpy(dst + at + 1, s, n);
    return at + 1 + (uint32_t)n;
}

/* ref: postSpots(), rtlsdr_ft8d.c:381-396 (receiver options template) and :398-423 (sender template).
 * Written out as the IPFIX fields they encode: {element id, length, enterprise number 0x0000768F}. */
static uint32_t rx_template(unsigned char *d) {
    static const uint16_t ids[3] = {0x8002, 0x8004, 0x8008};
    uint32_t n = 0;
    be16(d + n, 0x0003); n += 2;   /* set id: options template */
    be16(d + n, 0x0024); n += 2;   /* set length 36 */
    be16(d + n, 0x9992); n += 2;   /* template (link) id */
    be16(d + n, 3); n += 2;        /* field count */
    be16(d + n, 0); n += 2;        /* scope field count */
    for (int k = 0; k < 3; ++k) { be16(d + n, ids[k]); be16(d + n + 2, 0xFFFF); be32(d + n + 4, 0x0000768F); n += 8; }
    be16(d + n, 0); n += 2;        /* padding */
    return n;
}
static uint32_t tx_template(unsigned char *d) {
    static const uint16_t ids[6] = {0x8001, 0x8005, 0x8006, 0x800A, 0x8003, 0x800B};
    static const uint16_t lens[6] = {0xFFFF, 4, 1, 0xFFFF, 0xFFFF, 1};
    uint32_t n = 0;
    be16(d + n, 0x0002); n += 2;   /* set id: template */
    be16(d + n, 0x003C); n += 2;   /* set length 60 */
    be16(d + n, 0x9993); n += 2;
    be16(d + n, 7); n += 2;
    for (int k = 0; k < 6; ++k) { be16(d + n, ids[k]); be16(d + n + 2, lens[k]); be32(d + n + 4, 0x0000768F); n += 8; }
    be16(d + n, 0x0096); be16(d + n + 2, 4); n += 4;   /* dateTimeSeconds, 4 bytes, no enterprise number */
    return n;
}

/* ref: postSpots(), rtlsdr_ft8d.c:365-552 without the early return of :366.  Returns the datagram length. */
int orc_pskreporter_datagram(const orc_result_t *spots, uint32_t n_spots, const char *rcall, const char *rloc, uint32_t dial_freq,
                             const char *app_version, uint32_t unixtime, uint32_t sequence, uint32_t random_id, unsigned char *out) {
    unsigned char header[16], rxd[64], txd[64], rx[512], tx[1500];
    memset(rx, 0, sizeof(rx));
    memset(tx, 0, sizeof(tx));
    /* :438-450 */
    be16(header + 0, 0x000A);
    be32(header + 4, unixtime);
    be32(header + 8, sequence);
    be32(header + 12, random_id);
    const uint32_t rxd_n = rx_template(rxd), txd_n = tx_template(txd);
    /* :453-483 */
    uint32_t r = 0;
    be16(rx + r, 0x9992); r += 4;
    r = pstr(rx, r, rcall);
    r = pstr(rx, r, rloc);
    r = pstr(rx, r, app_version);
    if (r % 4) r += 4 - r % 4;
    /* :486-538 */
    uint32_t t = 0;
    be16(tx + t, 0x9993); t += 4;
    for (uint32_t i = 0; i < n_spots; ++i) {
        if (t > 1200) break;                                   /* :498 */
        t = pstr(tx, t, spots[i].call);
        be32(tx + t, (uint32_t)spots[i].freq + dial_freq); t += 4;
        tx[t] = (unsigned char)((int8_t)spots[i].snr - 20); t += 1;   /* :511 */
        t = pstr(tx, t, "FT8");
        t = pstr(tx, t, spots[i].loc);
        tx[t] = 1; t += 1;
        be32(tx + t, unixtime); t += 4;
    }
    if (t % 4) t += 4 - t % 4;
    /* :541-552 */
    const uint32_t full = 16 + rxd_n + txd_n + r + t;
    be16(rx + 2, r);
    be16(tx + 2, t);
    be16(header + 2, full);
    uint32_t n = 0;
    memcpy(out + n, header, 16); n += 16;
    memcpy(out + n, rxd, rxd_n); n += rxd_n;
    memcpy(out + n, txd, txd_n); n += txd_n;
    memcpy(out + n, rx, r); n += r;
    memcpy(out + n, tx, t); n += t;
    return (int)n;
}

/* ref: webClusterSpots(), rtlsdr_ft8d.c:603-606: the four strings of one spot's POST; out = 4 fields x 112 bytes */
void orc_webcluster_form(const orc_result_t *spot, const char *rcall, const char *rloc, uint32_t dial_freq, char *out) {
    char my[16], dx[12], fr[10], info[100];
    snprintf(my, sizeof(my), "%s", rcall);
    snprintf(fr, sizeof(fr), "%8f", (float)((uint32_t)spot->freq + dial_freq) / 1000.0f);
    snprintf(dx, sizeof(dx), "%s", spot->call);
    snprintf(info, sizeof(info), "M2M FT8 [%s - %s]", rloc, spot->loc);
    memset(out, 0, 4 * 112);
    strcpy(out + 0 * 112, my);
    strcpy(out + 1 * 112, dx);
    strcpy(out + 2 * 112, fr);
    strcpy(out + 3 * 112, info);
}

/* ref: printSpots(), rtlsdr_ft8d.c:635-663 (rx_state.gtm = gmtime of the slot time, :271).  Returns characters written. */
int orc_print_spots(const orc_result_t *spots, uint32_t n_spots, uint32_t dial_freq, uint32_t unixtime, char *out, size_t cap) {
    size_t n = 0;
    if (n_spots == 0) {
        time_t t = (time_t)unixtime;
        struct tm g;
        gmtime_r(&t, &g);
        return snprintf(out, cap, "No spot %04d-%02d-%02d %02d:%02dz\n", g.tm_year + 1900, g.tm_mon + 1, g.tm_mday, g.tm_hour, g.tm_min);
    }
    n += (size_t)snprintf(out + n, cap - n, "  Score     Freq       Call    Loc\n");
    for (uint32_t i = 0; i < n_spots && n < cap; ++i)
        n += (size_t)snprintf(out + n, cap - n, "     %2d %8d %10s %6s\n", spots[i].snr, (int)((uint32_t)spots[i].freq + dial_freq), spots[i].call, spots[i].loc);
    return (int)n;
}
