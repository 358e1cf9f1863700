// sidm.cu - placeholder entry points (implemented next)
#include "ctx.cuh"
using namespace b200;
extern "C" int b200_sidm(const int *, int, double, double, const b200_replay *) { return B200_ERR_STATE; }
extern "C" int b200_setup_nbr_sidm(const int *, int) { return B200_ERR_STATE; }
extern "C" int b200_sidm_ensure_neighbours(int, double, double, const b200_replay *) { return B200_ERR_STATE; }
extern "C" int b200_setup_smoothinglengths_sidm(int) { return B200_ERR_STATE; }
extern "C" int b200_compute_accelerations(int, const int *, int, double, double) { return B200_ERR_STATE; }
extern "C" int b200_ngb_treefind(const int *, int, int, float *) { return B200_ERR_STATE; }
extern "C" int b200_ngb_lists(const int *, int, int, int *, int *) { return B200_ERR_STATE; }
extern "C" int b200_sidm_debug(int, int *, double *, double *, int *) { return B200_ERR_STATE; }
extern "C" int b200_get_scatlog(b200_scatlog *, int, int *) { return B200_ERR_STATE; }
