// Drives csrc/sqlite_glue.cpp (sqlite3_vec_init + the vec_distance_* SQL scalars) through a minimal in-process stand-in for
// the SQLite C API: the glue is compiled against tests/host_cpp/mock_sqlite3/sqlite3.h and "SELECT vec_distance_l2(a, b)" is
// modelled by calling the registered xFunc with two value objects.  KATs: sqlite-vec-cpp tests/test_distances.cpp:20-53 and
// tests/unit/vector/sqlite_vec_c_api_smoke_catch2_test.cpp:19-75.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mock_sqlite3/sqlite3.h"

struct sqlite3 {
    std::map<std::string, void (*)(sqlite3_context*, int, sqlite3_value**)> fns;
};
struct sqlite3_value {
    int type = SQLITE_NULL;
    unsigned subtype = 0;
    std::vector<unsigned char> blob;
};
struct sqlite3_context {
    bool is_error = false;
    std::string error;
    double value = 0.0;
};
extern "C" {
int sqlite3_value_type(sqlite3_value* v) { return v->type; }
unsigned int sqlite3_value_subtype(sqlite3_value* v) { return v->subtype; }
const void* sqlite3_value_blob(sqlite3_value* v) { return v->blob.data(); }
int sqlite3_value_bytes(sqlite3_value* v) { return (int)v->blob.size(); }
void sqlite3_result_error(sqlite3_context* c, const char* m, int) { c->is_error = true; c->error = m; }
void sqlite3_result_double(sqlite3_context* c, double d) { c->is_error = false; c->value = d; }
char* sqlite3_mprintf(const char* fmt, ...) {
    char* buf = (char*)std::malloc(512);
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, 512, fmt, ap);
    va_end(ap);
    return buf;
}
int sqlite3_create_function_v2(sqlite3* db, const char* name, int nArg, int, void*, void (*xFunc)(sqlite3_context*, int, sqlite3_value**),
                               void (*)(sqlite3_context*, int, sqlite3_value**), void (*)(sqlite3_context*), void (*)(void*)) {
    if (nArg != 2 || !xFunc) return SQLITE_ERROR;
    db->fns[name] = xFunc;
    return SQLITE_OK;
}
int sqlite3_vec_init(sqlite3* db, char** pzErrMsg, const sqlite3_api_routines* pApi);
}

#define CHECK(x)                                                                          \
    do {                                                                                  \
        if (!(x)) {                                                                       \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #x);     \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

static sqlite3_value blob_of(const std::vector<float>& v, unsigned subtype = 0) {
    sqlite3_value x;
    x.type = SQLITE_BLOB;
    x.subtype = subtype;
    x.blob.resize(v.size() * 4);
    if (!v.empty()) std::memcpy(x.blob.data(), v.data(), v.size() * 4);
    return x;
}
static sqlite3_context call(sqlite3& db, const char* fn, sqlite3_value a, sqlite3_value b) {
    sqlite3_context ctx;
    sqlite3_value* argv[2] = {&a, &b};
    db.fns.at(fn)(&ctx, 2, argv);
    return ctx;
}

int main() {
    sqlite3 db;
    char* err = nullptr;
    CHECK(sqlite3_vec_init(nullptr, &err, nullptr) == SQLITE_ERROR && err && std::strstr(err, "database handle is null"));
    std::free(err);
    CHECK(sqlite3_vec_init(&db, &err, nullptr) == SQLITE_OK);
    CHECK(db.fns.count("vec_distance_l2") && db.fns.count("vec_distance_cosine") && db.fns.count("vec_distance_l1"));
    // test_distances.cpp:20-53
    auto r = call(db, "vec_distance_l2", blob_of({1, 2, 3, 4}), blob_of({2, 3, 4, 5}, 223));
    CHECK(!r.is_error && std::fabs(r.value - 2.0) < 1e-6);
    r = call(db, "vec_distance_l1", blob_of({1, 2, 3, 4}), blob_of({2, 3, 4, 5}));
    CHECK(!r.is_error && std::fabs(r.value - 4.0) < 1e-6);
    r = call(db, "vec_distance_cosine", blob_of({1, 0, 0}), blob_of({0, 1, 0}));
    CHECK(!r.is_error && std::fabs(r.value - 1.0) < 1e-6);
    r = call(db, "vec_distance_cosine", blob_of({1, 2, 3}), blob_of({2, 4, 6}));
    CHECK(!r.is_error && std::fabs(r.value) < 1e-6);
    std::vector<float> big_a(768), big_b(768);
    for (int i = 0; i < 768; ++i) { big_a[i] = std::sin(0.1f * i); big_b[i] = std::cos(0.07f * i); }
    double want = 0;
    for (int i = 0; i < 768; ++i) want += (double)(big_a[i] - big_b[i]) * (big_a[i] - big_b[i]);
    r = call(db, "vec_distance_l2", blob_of(big_a), blob_of(big_b));
    CHECK(!r.is_error && std::fabs(r.value - std::sqrt(want)) < 1e-3);
    // error contract (functions.hpp:31-60, 86-118)
    r = call(db, "vec_distance_l2", blob_of({1, 2, 3}), blob_of({1, 2}));
    CHECK(r.is_error && r.error == "Dimension mismatch: expected 3, got 2");
    sqlite3_value nul;
    r = call(db, "vec_distance_l2", nul, blob_of({1, 2}));
    CHECK(r.is_error && r.error == "Vector value is NULL");
    sqlite3_value odd = blob_of({1, 2});
    odd.blob.pop_back();
    r = call(db, "vec_distance_cosine", odd, blob_of({1, 2}));
    CHECK(r.is_error && r.error == "Blob size not aligned to element size");
    r = call(db, "vec_distance_l2", blob_of({1, 2}, 224), blob_of({1, 2}, 224));
    CHECK(r.is_error && r.error == "Cannot calculate L2 distance between bitvectors");
    r = call(db, "vec_distance_l2", blob_of({1, 2}, 225), blob_of({1, 2}, 223));
    CHECK(r.is_error && r.error == "Vector element types must match");
    {
        sqlite3_context ctx;
        sqlite3_value a = blob_of({1, 2});
        sqlite3_value* argv1[1] = {&a};
        db.fns.at("vec_distance_l1")(&ctx, 1, argv1);
        CHECK(ctx.is_error && ctx.error == "vec_distance_l1 requires exactly 2 arguments");
    }
    std::puts("GLUE OK");
    return 0;
}
