// fkm_host.h — host-side helpers of the library (FASTA packing, bin-file writer).
#pragma once
#include <cstdint>
#include <vector>

int fkm_set_error(int code, const char* fmt, ...);
int fkm_read_file(const char* path, std::vector<uint8_t>& out);
// writes <out_dir>/bin<id> for every non-empty bin; sorted!=0 appends the "EOF" trailer (SBKC:598-606)
int fkm_write_bins(const char* out_dir, int32_t B, int32_t k, int sorted, const uint64_t* out_base,
                   const uint64_t* hi, const uint64_t* lo, const uint32_t* cnt);
int fkm_make_dirs(const char* dir);      // mkdir -p
