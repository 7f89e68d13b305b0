/* Stand-in for <curl/curl.h>: the reference's web-cluster POST is out of scope
 * (SURVEY.md §2 row 20); these stubs only let the daemon source compile as an oracle. */
#ifndef ORACLE_SHIM_CURL_H
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
static inline int curl_global_init(long f) { (void)f; return 0; }
static inline int curl_formadd(struct curl_httppost **p, struct curl_httppost **l, ...) { (void)p; (void)l; return 0; }
static inline CURL *curl_easy_init(void) { return 0; }
static inline int curl_easy_setopt(CURL *c, int o, ...) { (void)c; (void)o; return 0; }
static inline CURLcode curl_easy_perform(CURL *c) { (void)c; return 1; }
static inline const char *curl_easy_strerror(CURLcode c) { (void)c; return "curl stub"; }
static inline void curl_easy_cleanup(CURL *c) { (void)c; }
static inline void curl_formfree(struct curl_httppost *p) { (void)p; }
#endif
