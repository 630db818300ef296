// TEST INFRASTRUCTURE ONLY (oracle/): force-included (-include) when compiling a reference
// testbench.  The testbenches open their operand with an absolute path on the author's machine
// ("/curr/jieliu15/HLS_design/<Algo>/op<N>.bin"); map that prefix to $HLS_REF_OPERANDS so the
// unmodified testbench reads the same op<N>.bin from wherever the operand files live here.
#ifndef ORACLE_SHIM_TB_FOPEN_REDIRECT_H
#define ORACLE_SHIM_TB_FOPEN_REDIRECT_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
static inline FILE* oracle_tb_fopen(const char* path, const char* mode) {
    static const char kPrefix[] = "/curr/jieliu15/HLS_design/";
    const char* root = std::getenv("HLS_REF_OPERANDS");
    if (root && std::strncmp(path, kPrefix, sizeof(kPrefix) - 1) == 0) {
        std::string p = std::string(root) + "/" + (path + sizeof(kPrefix) - 1);
        return std::fopen(p.c_str(), mode);
    }
    return std::fopen(path, mode);
}
#define fopen(p, m) oracle_tb_fopen((p), (m))
#endif
