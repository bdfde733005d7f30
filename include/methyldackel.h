/*
 * methyldackel.h — the sub-command entry points of the reference, with the reference's own signatures
 * (main.c:17-20: `int extract_main(int argc, char *argv[])` etc., which main.c:52-59 dispatches to and which
 * libMethylDackel.a exports, Makefile:22-26), bound to the B200 back end (libmdgpu).  A program that links
 * libMethylDackel.a and calls extract_main()/mbias_main()/perRead_main() links lib/libMethylDackel.so instead
 * and is unchanged.  argv[0] is the sub-command name, as in the reference.
 *
 * The CUDA device is taken from the environment variable MD_DEVICE (default 0).  Without a usable CUDA device the
 * calls fail with an error message and a non-zero return value; there is no CPU fallback.
 */
#ifndef METHYLDACKEL_DROPIN_H
#define METHYLDACKEL_DROPIN_H
#include "mdhost.h"
#ifdef __cplusplus
extern "C" {
#endif

int extract_main(int argc, char *argv[]);   /* extract.c:706 */
int mbias_main(int argc, char *argv[]);     /* MBias.c:304   */
int perRead_main(int argc, char *argv[]);   /* perRead.c:305 */

/* The back-end table these use: every slot bound to libmdgpu on CUDA device `device` (what mdh_*_main take). */
const mdh_backend *mdh_gpu_backend(int device);

#ifdef __cplusplus
}
#endif
#endif
