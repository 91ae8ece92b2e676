// stages_stub.cu -- entry points declared in include/sindyn.h whose kernels are not built yet.
// They fail loudly (SINDYN_ERR_STATE + message); there is no CPU fallback.
#include "ctx.cuh"

#define NOT_YET(h, name) do { if (!(h)) return SINDYN_ERR_INVALID; (h)->err = name ": stage not built yet"; return SINDYN_ERR_STATE; } while (0)



struct sindyn_orb : sindyn_base {};
extern "C" int sindyn_orb_create(int, float, int, int, int, int, int, int, sindyn_orb_handle *) { return SINDYN_ERR_STATE; }
extern "C" int sindyn_orb_destroy(sindyn_orb_handle) { return SINDYN_ERR_STATE; }
extern "C" int sindyn_orb_extract(sindyn_orb_handle, const uint8_t *, size_t, const uint8_t *, size_t, sindyn_keypoint *, uint8_t *, int, int *) { return SINDYN_ERR_STATE; }
extern "C" int sindyn_orb_get_pyramid_level(sindyn_orb_handle, int, uint8_t *, int *, int *) { return SINDYN_ERR_STATE; }
extern "C" int sindyn_orb_set_stream(sindyn_orb_handle, void *) { return SINDYN_ERR_STATE; }
extern "C" unsigned long long sindyn_orb_launch_count(sindyn_orb_handle) { return 0; }
extern "C" const char *sindyn_orb_last_error(sindyn_orb_handle) { return "orb: not built yet"; }
