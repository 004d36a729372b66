/* TEST SUPPORT (LD_PRELOAD): the reference's eval drivers open their datasets under the
 * hard-coded prefix /workspace/data/ (Auncel/eval/bound.cpp:155-200).  To run such a driver
 * UNMODIFIED on synthetic files, fopen() calls below that prefix are redirected to
 * $AUNCEL_DATA_ROOT.  Nothing else is touched. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char PREFIX[] = "/workspace/data/";

static const char* remap(const char* path, char* buf, size_t cap) {
    const char* root = getenv("AUNCEL_DATA_ROOT");
    if (!root || !path || strncmp(path, PREFIX, sizeof(PREFIX) - 1) != 0) return path;
    snprintf(buf, cap, "%s/%s", root, path + sizeof(PREFIX) - 1);
    return buf;
}

FILE* fopen(const char* path, const char* mode) {
    static FILE* (*real)(const char*, const char*) = NULL;
    if (!real) real = (FILE * (*)(const char*, const char*)) dlsym(RTLD_NEXT, "fopen");
    char buf[4096];
    return real(remap(path, buf, sizeof(buf)), mode);
}

FILE* fopen64(const char* path, const char* mode) {
    static FILE* (*real)(const char*, const char*) = NULL;
    if (!real) real = (FILE * (*)(const char*, const char*)) dlsym(RTLD_NEXT, "fopen64");
    char buf[4096];
    return real(remap(path, buf, sizeof(buf)), mode);
}
