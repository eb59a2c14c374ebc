/* TEST INFRASTRUCTURE -- not product code.
 * Minimal stand-in for <curl/curl.h> so that the UNMODIFIED reference sources
 * (/root/reference/gps.c, almanac.c) compile in a container without libcurl.
 * Only the identifiers those two files use are provided (gps.c:2367-2463,
 * almanac.c:191-215). All transfers report CURLE_GOT_NOTHING: the oracle never
 * downloads anything (no network). */
#ifndef ORACLE_STUB_CURL_H
#define ORACLE_STUB_CURL_H
#include <stddef.h>

typedef void CURL;
typedef enum {
    CURLE_OK = 0,
    CURLE_READ_ERROR = 26,
    CURLE_GOT_NOTHING = 52,
    CURLE_REMOTE_FILE_NOT_FOUND = 78
} CURLcode;
typedef enum {
    CURLOPT_URL = 10002,
    CURLOPT_WRITEFUNCTION = 20011,
    CURLOPT_WRITEDATA = 10001,
    CURLOPT_USE_SSL = 119,
    CURLOPT_VERBOSE = 41,
    CURLOPT_USERPWD = 10005
} CURLoption;
#define CURLUSESSL_NONE 0L
#define CURL_GLOBAL_DEFAULT 3L

static inline CURLcode curl_global_init(long f) { (void) f; return CURLE_OK; }
static inline void curl_global_cleanup(void) {}
static inline CURL *curl_easy_init(void) { return NULL; }
#define curl_easy_setopt(h, o, v) ((void) (h), (void) (o), (void) (v), CURLE_OK)
static inline CURLcode curl_easy_perform(CURL *h) { (void) h; return CURLE_GOT_NOTHING; }
static inline void curl_easy_cleanup(CURL *h) { (void) h; }
#endif
