// Reader for the PJB200T1 mechanism-table container written by pyjac_b200/blob.py.
//
//   char magic[8] = "PJB200T1"; int64 n_entries;
//   entry[n]: char name[24]; int32 dtype (0 f64, 1 i32, 2 u16); int32 pad; int64 count; int64 offset
//   payload (each array 16-byte aligned, offsets from the start of the blob)
#pragma once
#include <cstdint>
#include <cstring>

namespace pjt {

// versions this library was written for: table set, p5_* schedule tables, p6_* record streams
// (pyjac_b200/tables.py SCHEMA_VERSION, plan.py / plan6.py PLAN_VERSION)
enum : int32_t { SCHEMA_VERSION = 2, PLAN5_VERSION = 3, PLAN6_VERSION = 4 };

struct Entry {
    char name[24];
    int32_t dtype, pad;
    int64_t count, offset;
};

inline size_t elem_size(int dtype) { return dtype == 0 ? 8 : (dtype == 1 ? 4 : 2); }

inline bool valid(const void* blob, size_t len)
{
    if (!blob || len < 16 || std::memcmp(blob, "PJB200T1", 8) != 0) return false;
    const char* b = static_cast<const char*>(blob);
    int64_t n;
    std::memcpy(&n, b + 8, 8);
    if (n < 0 || 16 + (size_t)n * sizeof(Entry) > len) return false;
    const Entry* e = reinterpret_cast<const Entry*>(b + 16);
    for (int64_t k = 0; k < n; ++k) {
        if (e[k].dtype < 0 || e[k].dtype > 2 || e[k].count < 0 || e[k].offset < 0) return false;
        if ((size_t)e[k].offset + (size_t)e[k].count * elem_size(e[k].dtype) > len) return false;
    }
    return true;
}

inline const Entry* find(const void* blob, const char* name)
{
    const char* b = static_cast<const char*>(blob);
    int64_t n;
    std::memcpy(&n, b + 8, 8);
    const Entry* e = reinterpret_cast<const Entry*>(b + 16);
    for (int64_t k = 0; k < n; ++k)
        if (std::strncmp(e[k].name, name, 24) == 0) return &e[k];
    return nullptr;
}

}  // namespace pjt
