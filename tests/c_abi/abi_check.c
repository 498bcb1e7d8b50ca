/* C99 consumer of include/clm_b200.h: the header must be plain C (a Julia ccall / cgo / JNI binding sees exactly this
 * ABI), the library must load with dlopen and answer the calls that need no device.  Used by tests/test_host_cpu.py. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "clm_b200.h"

typedef int (*version_fn)(void);
typedef int (*create_fn)(clm_handle**, int, int, int, int);
typedef const char* (*error_fn)(clm_handle*);
typedef int (*destroy_fn)(clm_handle*);

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: abi_check <libclm_b200.so>\n"); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
    version_fn version;
    create_fn create;
    error_fn last_error;
    destroy_fn destroy;
    *(void**)(&version) = dlsym(lib, "clm_version");      /* the POSIX idiom for object -> function pointer */
    *(void**)(&create) = dlsym(lib, "clm_create");
    *(void**)(&last_error) = dlsym(lib, "clm_last_error");
    *(void**)(&destroy) = dlsym(lib, "clm_destroy");
    if (!version || !create || !last_error || !destroy) { fprintf(stderr, "missing symbol\n"); return 4; }
    printf("version %d\n", version());
    printf("sizeof clm_box_info %d clm_stats %d clm_custom_info %d\n", (int)sizeof(clm_box_info), (int)sizeof(clm_stats), (int)sizeof(clm_custom_info));
    clm_handle* h = NULL;
    int rc = create(&h, 4, CLM_F32, 0, 1);            /* argument validation happens before any device work */
    printf("create(dim=4) -> %d: %s\n", rc, last_error(NULL));
    if (rc != CLM_ERR_DIMENSION) return 5;
    rc = create(&h, 3, CLM_F64, 0, 1);
    printf("create(dim=3) -> %d%s%s\n", rc, rc ? ": " : "", rc ? last_error(NULL) : "");
    if (rc == CLM_OK) destroy(h);                       /* a device is present */
    else if (rc != CLM_ERR_CUDA) return 6;              /* no device: the product fails loudly, there is no CPU fallback */
    return 0;
}
