// Test-only stand-in for <rocksdb/db.h>: index_builder/build.cpp (reference) stores every vector in RocksDB
// (build.cpp:127-142), which is storage, out of scope and not installed here.  This stub lets build.cpp compile
// UNCHANGED against the drop-in hnswlib header; Put() discards its arguments.
#pragma once
#include <string>
namespace rocksdb {
struct Options { bool create_if_missing = false; };
struct WriteOptions {};
struct Status {
    bool ok() const { return true; }
    std::string ToString() const { return "OK"; }
};
class DB {
 public:
    static Status Open(const Options &, const std::string &, DB **db) { *db = new DB(); return Status(); }
    Status Put(const WriteOptions &, const std::string &, const std::string &) { return Status(); }
};
}  // namespace rocksdb
