// Error reporting, versioning and launch accounting for the C ABI.
#include "common.cuh"

static thread_local std::string t_last_error;
unsigned long long g_dm_launches = 0;

void dm_set_error(const std::string& msg) { t_last_error = msg; }

extern "C" int dm_abi_version(void) { return DM_ABI_VERSION; }
extern "C" const char* dm_last_error(void) { return t_last_error.c_str(); }
extern "C" uint64_t dm_launch_count(void) { return g_dm_launches; }
