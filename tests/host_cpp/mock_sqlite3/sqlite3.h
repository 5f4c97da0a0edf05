/* A stand-in for <sqlite3.h> with only what csrc/sqlite_glue.cpp uses, so that the SQL glue can be compiled and driven by
 * tests/host_cpp/test_sqlite_glue.cpp in an image that has no SQLite development header.  Values mirror sqlite3.h. */
#ifndef YAMS_B200_MOCK_SQLITE3_H
#define YAMS_B200_MOCK_SQLITE3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sqlite3 sqlite3;
typedef struct sqlite3_context sqlite3_context;
typedef struct sqlite3_value sqlite3_value;
typedef struct sqlite3_api_routines sqlite3_api_routines;
#define SQLITE_OK 0
#define SQLITE_ERROR 1
#define SQLITE_BLOB 4
#define SQLITE_NULL 5
#define SQLITE_UTF8 1
#define SQLITE_DETERMINISTIC 0x000000800
#define SQLITE_SUBTYPE 0x000100000
int sqlite3_value_type(sqlite3_value*);
unsigned int sqlite3_value_subtype(sqlite3_value*);
const void* sqlite3_value_blob(sqlite3_value*);
int sqlite3_value_bytes(sqlite3_value*);
void sqlite3_result_error(sqlite3_context*, const char*, int);
void sqlite3_result_double(sqlite3_context*, double);
char* sqlite3_mprintf(const char*, ...);
int sqlite3_create_function_v2(sqlite3* db, const char* zFunctionName, int nArg, int eTextRep, void* pApp,
                               void (*xFunc)(sqlite3_context*, int, sqlite3_value**),
                               void (*xStep)(sqlite3_context*, int, sqlite3_value**), void (*xFinal)(sqlite3_context*),
                               void (*xDestroy)(void*));
#ifdef __cplusplus
}
#endif
#endif
