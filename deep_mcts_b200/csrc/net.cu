// placeholder until the network kernels land (next commit)
#include "common.cuh"
#define DM_NET_STUB(name, ...) extern "C" int name(__VA_ARGS__) { dm_set_error(#name ": network kernels not built yet"); return DM_ERR_STATE; }
DM_NET_STUB(dm_net_create, const dm_net_config*, dm_net**)
DM_NET_STUB(dm_net_destroy, dm_net*)
DM_NET_STUB(dm_net_set_tensor, dm_net*, const char*, const float*, int64_t)
DM_NET_STUB(dm_net_finalize, dm_net*, void*)
DM_NET_STUB(dm_net_evaluate, dm_net*, int, int, const int32_t*, const uint64_t*, const uint64_t*, const uint8_t*, float*, float*, void*)
DM_NET_STUB(dm_net_forward, dm_net*, int, int, const uint64_t*, const uint64_t*, const uint8_t*, float*, float*, void*)
DM_NET_STUB(dm_net_rollout_greedy, dm_net*, int, int, const uint64_t*, const uint64_t*, const uint8_t*, float*, void*)
