/* oracle/ref_report_harness.c -- the reference daemon's reporting functions (postSpots, webClusterSpots,
 * printSpots: rtlsdr_ft8d.c:365-663) made callable so their OUTPUT BYTES can be captured.
 * TEST INFRASTRUCTURE ONLY: nothing under oracle/ is imported, linked or executed by the product path.
 *
 * postSpots() is dead code upstream: its first statement is `return;` (rtlsdr_ft8d.c:366) and, past it, the
 * connect loop starts from an uninitialised pointer (`for (rp = res; ...`, :560).  oracle/Makefile therefore
 * compiles a TEMP COPY of the daemon source (never stored in this repository) with exactly those two lines
 * edited by sed: the early `return;` deleted and the loop seeded from getaddrinfo()'s result.  Everything that
 * builds the datagram (:374-552) is the reference's own code.
 *
 * Capture points (no network, no terminal):
 *   - send/socket/connect/close/getaddrinfo/freeaddrinfo/time are redirected with -Wl,--wrap to the stubs
 *     below; __wrap_send keeps the datagram, __wrap_time returns the time the test asked for;
 *   - printf is renamed while the daemon source is included, so printSpots' lines land in a buffer;
 *   - the curl form calls (shims/curl/curl.h is replaced here) record the four fields of every POST.
 */
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <sys/types.h>
#include <sys/socket.h>
#include <netdb.h>
#include <time.h>

/* ---- curl stand-in that records curl_formadd(name, contents) pairs ------------------------------------ */
#define ORACLE_SHIM_CURL_H
typedef void CURL;
typedef int CURLcode;
struct curl_httppost { int unused; };
#define CURLE_OK 0
#define CURL_GLOBAL_ALL 3
#define CURLFORM_COPYNAME 1
#define CURLFORM_COPYCONTENTS 4
#define CURLFORM_END 17
#define CURLOPT_URL 10002
#define CURLOPT_HTTPPOST 10024
#define REP_MAX_FIELDS 1024
static char g_form[REP_MAX_FIELDS][2][112];
static int g_form_n;
static int curl_global_init(long f) { (void)f; return 0; }
static int curl_formadd(struct curl_httppost **p, struct curl_httppost **l, ...) {
    va_list ap;
    (void)p; (void)l;
    va_start(ap, l);
    int o1 = va_arg(ap, int); const char *name = va_arg(ap, const char *);
    int o2 = va_arg(ap, int); const char *val = va_arg(ap, const char *);
    va_end(ap);
    if (o1 == CURLFORM_COPYNAME && o2 == CURLFORM_COPYCONTENTS && g_form_n < REP_MAX_FIELDS) {
        snprintf(g_form[g_form_n][0], sizeof(g_form[0][0]), "%s", name);
        snprintf(g_form[g_form_n][1], sizeof(g_form[0][1]), "%s", val);
        ++g_form_n;
    }
    return 0;
}
static CURL *curl_easy_init(void) { return 0; }
static int curl_easy_setopt(CURL *c, int o, ...) { (void)c; (void)o; return 0; }
static CURLcode curl_easy_perform(CURL *c) { (void)c; return 1; }
static const char *curl_easy_strerror(CURLcode c) { (void)c; return "curl stub"; }
static void curl_easy_cleanup(CURL *c) { (void)c; }
static void curl_formfree(struct curl_httppost *p) { (void)p; }

/* ---- stdout capture -------------------------------------------------------------------------------- */
static char g_out[65536];
static size_t g_out_n;
static int tap_printf(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(g_out + g_out_n, sizeof(g_out) - g_out_n, fmt, ap);
    va_end(ap);
    if (n > 0) g_out_n += (size_t)n < sizeof(g_out) - g_out_n ? (size_t)n : sizeof(g_out) - g_out_n - 1;
    return n;
}

#define main ref_daemon_main
#define printf tap_printf
#include "rtlsdr_ft8d.c" /* the temp copy: -I$(TMP) precedes -I$(REF) */
#undef printf
#undef main

/* ---- network stand-ins (-Wl,--wrap=...) -------------------------------------------------------------- */
static unsigned char g_pkt[4096];
static int g_pkt_len = -1;
static uint32_t g_fake_time;
time_t __real_time(time_t *t);
int __real_close(int fd);
#define REP_FD 7777

ssize_t __wrap_send(int fd, const void *buf, size_t n, int flags) {
    (void)fd; (void)flags;
    g_pkt_len = (int)(n < sizeof(g_pkt) ? n : sizeof(g_pkt));
    memcpy(g_pkt, buf, (size_t)g_pkt_len);
    return (ssize_t)n;
}
int __wrap_getaddrinfo(const char *node, const char *service, const struct addrinfo *hints, struct addrinfo **res) {
    static struct addrinfo ai;
    static struct sockaddr sa;
    (void)node; (void)service; (void)hints;
    memset(&ai, 0, sizeof(ai));
    ai.ai_family = AF_INET; ai.ai_socktype = SOCK_DGRAM; ai.ai_addr = &sa; ai.ai_addrlen = sizeof(sa);
    *res = &ai;
    return 0;
}
void __wrap_freeaddrinfo(struct addrinfo *res) { (void)res; }
int __wrap_socket(int d, int t, int p) { (void)d; (void)t; (void)p; return REP_FD; }
int __wrap_connect(int fd, const struct sockaddr *a, socklen_t l) { (void)fd; (void)a; (void)l; return 0; }
int __wrap_close(int fd) { return fd == REP_FD ? 0 : __real_close(fd); }
time_t __wrap_time(time_t *t) {
    if (!g_fake_time) return __real_time(t);
    if (t) *t = (time_t)g_fake_time;
    return (time_t)g_fake_time;
}

/* ---- entry points ------------------------------------------------------------------------------------ */
int ref_report_max_messages(void) { return (int)(sizeof(dec_results) / sizeof(dec_results[0])); }
const char *ref_report_app_version(void) { return pskreporter_app_version; }

static void load(const struct decoder_results *spots, uint32_t n, const char *rcall, const char *rloc, uint32_t dial, uint32_t unixtime) {
    memset(dec_results, 0, sizeof(dec_results));
    memcpy(dec_results, spots, sizeof(struct decoder_results) * n);
    memset(&dec_options, 0, sizeof(dec_options));
    dec_options.freq = dial;
    snprintf(dec_options.rcall, sizeof(dec_options.rcall), "%s", rcall);
    snprintf(dec_options.rloc, sizeof(dec_options.rloc), "%s", rloc);
    rx_options.noreport = false;
    g_fake_time = unixtime;
}

/* postSpots(): returns the datagram length (or -1 when nothing was sent); bytes 12..15 hold the process-wide random id */
int ref_post_spots(const struct decoder_results *spots, uint32_t n, const char *rcall, const char *rloc, uint32_t dial, uint32_t unixtime,
                   unsigned char *out, int cap) {
    load(spots, n, rcall, rloc, dial, unixtime);
    g_pkt_len = -1;
    postSpots(n);
    g_fake_time = 0;
    if (g_pkt_len > 0) memcpy(out, g_pkt, (size_t)(g_pkt_len < cap ? g_pkt_len : cap));
    return g_pkt_len;
}

/* printSpots(): returns the number of characters written to `out` */
int ref_print_spots(const struct decoder_results *spots, uint32_t n, uint32_t dial, uint32_t unixtime, char *out, int cap) {
    load(spots, n, "", "", dial, unixtime);
    time_t t = (time_t)unixtime;
    rx_state.gtm = gmtime(&t);
    g_out_n = 0; g_out[0] = 0;
    printSpots(n);
    g_fake_time = 0;
    snprintf(out, (size_t)cap, "%s", g_out);
    return (int)g_out_n;
}

/* webClusterSpots(): returns the number of (name, value) pairs recorded; pair k at out[k*2*112 ...] */
int ref_webcluster_spots(const struct decoder_results *spots, uint32_t n, const char *rcall, const char *rloc, uint32_t dial, char *out, int cap_pairs) {
    load(spots, n, rcall, rloc, dial, 0);
    g_form_n = 0;
    webClusterSpots(n);
    int k = g_form_n < cap_pairs ? g_form_n : cap_pairs;
    memcpy(out, g_form, (size_t)k * sizeof(g_form[0]));
    return g_form_n;
}
