// host_util.hpp -- small helpers shared by the host drivers (bc_run, bc_runner).
#pragma once
#include <string>
#include <sys/stat.h>

namespace {

// mkdir -p
void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty()) mkdir(cur.c_str(), 0777);
        }
        if (i < path.size()) cur += path[i];
    }
}

}  // namespace
