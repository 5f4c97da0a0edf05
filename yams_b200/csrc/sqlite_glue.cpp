// sqlite_glue.cpp -- the SQL face of the sqlite-vec-cpp operator surface, served by libyams_b200.
//
// Replaces third_party/sqlite-vec-cpp/src/sqlite_vec_c_api.cpp:15-55 (sqlite3_vec_init) and the scalar functions of
// include/sqlite-vec-cpp/sqlite/functions.hpp:76-278 as registered by sqlite/registration.hpp:42-58:
//     vec_distance_l2(a, b), vec_distance_cosine(a, b), vec_distance_l1(a, b)
// Argument contract kept from the reference: exactly two arguments; BLOBs; element type from the SQLite subtype
// (223 float32, 224 bit, 225 int8; 0 = stored blob = float32, functions.hpp:18-22,62-72); NULL / non-BLOB / size not a
// multiple of the element / misaligned / different dimensions are SQL errors with the reference's messages.  Only float32
// vectors are computed here (the device kernels of csrc/ref_order.cuh); int8 and bit vectors report an error that names
// sqlite-vec-cpp, which a build can still load beside this library for them.
//
// This file needs <sqlite3.h>.  The image this library is developed in has only the runtime libsqlite3.so.0, so the
// Makefile compiles it only when the header is found (`make SQLITE_GLUE=1`, or automatically if `sqlite3.h` is on the
// include path) -- the CMake equivalent is `find_package(SQLite3)` guarding this one source.  The vec0 virtual-table module stays in
// sqlite-vec-cpp; its exact plan calls yams_b200_vec0_exact.
#include <sqlite3.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/yams_b200.h"

namespace {

enum ElemType { kFloat32 = 223, kBit = 224, kInt8 = 225 };

ElemType elem_type_of(sqlite3_value* v) {
    const int st = (int)sqlite3_value_subtype(v);
    if (st == kInt8) return kInt8;
    if (st == kBit) return kBit;
    return kFloat32;   // 223, or 0 for blobs read back from a column (subtypes are not persisted)
}

// functions.hpp:31-60 extract_vector_from_value<float>
const char* extract_f32(sqlite3_value* v, const float** data, size_t* bytes) {
    if (sqlite3_value_type(v) == SQLITE_NULL) return "Vector value is NULL";
    if (sqlite3_value_type(v) != SQLITE_BLOB) return "Vector value must be BLOB type";
    const void* p = sqlite3_value_blob(v);
    const size_t n = (size_t)sqlite3_value_bytes(v);
    if (n % sizeof(float) != 0) return "Blob size not aligned to element size";
    if (reinterpret_cast<uintptr_t>(p) % alignof(float) != 0) return "Blob data is not aligned for requested vector type";
    *data = static_cast<const float*>(p);
    *bytes = n;
    return nullptr;
}

typedef int (*PairFn)(const void*, size_t, const void*, size_t, float*);

void distance_impl(sqlite3_context* ctx, int argc, sqlite3_value** argv, const char* name, const char* what, PairFn fn) {
    if (argc != 2) {
        sqlite3_result_error(ctx, (std::string(name) + " requires exactly 2 arguments").c_str(), -1);
        return;
    }
    const ElemType ta = elem_type_of(argv[0]), tb = elem_type_of(argv[1]);
    if (ta != tb) {
        sqlite3_result_error(ctx, "Vector element types must match", -1);
        return;
    }
    if (ta == kBit) {
        sqlite3_result_error(ctx, (std::string("Cannot calculate ") + what + " distance between bitvectors").c_str(), -1);
        return;
    }
    if (ta == kInt8) {
        sqlite3_result_error(ctx, (std::string(name) + ": int8 vectors are not served by yams_b200; register sqlite-vec-cpp for them").c_str(), -1);
        return;
    }
    const float *a = nullptr, *b = nullptr;
    size_t na = 0, nb = 0;
    const char* ea = extract_f32(argv[0], &a, &na);
    const char* eb = ea ? nullptr : extract_f32(argv[1], &b, &nb);
    if (ea || eb) {
        sqlite3_result_error(ctx, ea ? ea : eb, -1);
        return;
    }
    if (na != nb) {   // utils/error.hpp:162-166
        sqlite3_result_error(ctx, ("Dimension mismatch: expected " + std::to_string(na / 4) + ", got " + std::to_string(nb / 4)).c_str(), -1);
        return;
    }
    float dist = 0.f;
    if (na == 0) {   // an empty pair: every reference loop runs zero times
        sqlite3_result_double(ctx, fn == sqlite3_vec_distance_cosine ? 1.0 : 0.0);
        return;
    }
    if (fn(a, na, b, nb, &dist) != 0) {
        sqlite3_result_error(ctx, (std::string(name) + ": device evaluation failed: " + yams_b200_last_error()).c_str(), -1);
        return;
    }
    sqlite3_result_double(ctx, (double)dist);   // functions.hpp:124 result_double(static_cast<double>(dist))
}

void vec_distance_l2(sqlite3_context* c, int n, sqlite3_value** v) { distance_impl(c, n, v, "vec_distance_l2", "L2", sqlite3_vec_distance_l2); }
void vec_distance_cosine(sqlite3_context* c, int n, sqlite3_value** v) { distance_impl(c, n, v, "vec_distance_cosine", "cosine", sqlite3_vec_distance_cosine); }
void vec_distance_l1(sqlite3_context* c, int n, sqlite3_value** v) { distance_impl(c, n, v, "vec_distance_l1", "L1", yams_b200_vec_distance_l1); }

}  // namespace

extern "C" YAMS_B200_API int sqlite3_vec_init(sqlite3* db, char** pzErrMsg, const sqlite3_api_routines* pApi) {
    (void)pApi;   // direct sqlite3.h, like sqlite_vec_c_api.cpp:16
    if (!db) {
        if (pzErrMsg) *pzErrMsg = sqlite3_mprintf("Failed to initialize sqlite-vec: %s", "database handle is null");
        return SQLITE_ERROR;
    }
    if (yams_plugin_init(nullptr, nullptr) != YAMS_PLUGIN_OK) {
        if (pzErrMsg) *pzErrMsg = sqlite3_mprintf("Failed to initialize sqlite-vec: %s", yams_b200_last_error());
        return SQLITE_ERROR;
    }
    const int flags = SQLITE_UTF8 | SQLITE_DETERMINISTIC | SQLITE_SUBTYPE;   // registration.hpp:31 kVecReadFlags
    struct { const char* name; void (*fn)(sqlite3_context*, int, sqlite3_value**); } fns[] = {
        {"vec_distance_l2", vec_distance_l2}, {"vec_distance_l1", vec_distance_l1}, {"vec_distance_cosine", vec_distance_cosine}};
    for (auto& f : fns) {
        const int rc = sqlite3_create_function_v2(db, f.name, 2, flags, nullptr, f.fn, nullptr, nullptr, nullptr);
        if (rc != SQLITE_OK) {
            if (pzErrMsg) *pzErrMsg = sqlite3_mprintf("Failed to initialize sqlite-vec: Failed to register %s", f.name);
            return SQLITE_ERROR;
        }
    }
    return SQLITE_OK;
}
