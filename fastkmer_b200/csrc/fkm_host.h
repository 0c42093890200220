// fkm_host.h — host-side helpers of the library.
#pragma once
#include <cstdint>

int fkm_set_error(int code, const char* fmt, ...);
int fkm_make_dirs(const char* dir);      // mkdir -p
