/* Plain-C consumer of the drop-in boundary: proves include/strotss_b200.h is a C header (no C++/torch types) and that
 * the shared library links and loads on its own.  Built and run by tests/test_abi.py with gcc (no GPU needed: without a
 * device strotss_create must fail loudly and still hand back a handle whose error text can be read). */
#include <stdio.h>
#include <string.h>
#include "strotss_b200.h"

int main(void) {
    strotss_handle h = NULL;
    const char* v = strotss_version();
    if (!v || !strstr(v, "sm_100a")) { printf("bad version string\n"); return 2; }
    int rc = strotss_create(0, &h);
    printf("version=%s create_rc=%d phases=%d\n", v, rc, strotss_profile_num_phases());
    if (rc != STROTSS_OK) {
        if (!h) { printf("no handle returned on failure\n"); return 3; }
        printf("error=%s\n", strotss_last_error(h));
        if (strlen(strotss_last_error(h)) == 0) return 4;
    } else {
        /* with a GPU: an evaluation before a style target is set is a state error, not a crash */
        float dummy[STROTSS_NUM_SCALARS];
        rc = strotss_eval(h, dummy, 2179, dummy, 2179, 1, 16.0f, dummy, NULL, 0, NULL, NULL, NULL);
        if (rc != STROTSS_ERR_STATE) { printf("expected STROTSS_ERR_STATE, got %d\n", rc); return 5; }
    }
    strotss_destroy(h);
    printf("ok\n");
    return 0;
}
