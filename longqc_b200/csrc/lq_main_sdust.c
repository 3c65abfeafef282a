/* the `sdust` executable: everything lives in liblqcov.so */
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include "lqcov.h"
/* as lq_main_cov.c: stdout and the exit status are the contract; no orderly CUDA teardown unless LQCOV_FAST_EXIT=0 */
int main(int argc, char **argv)
{
    setenv("LQCOV_FAST_EXIT", "1", 0);                  /* tells the library not to tear its contexts down either */
    const int rc = lqcov_sdust_main(argc, argv);
    const char *fe = getenv("LQCOV_FAST_EXIT");
    if (fe && fe[0] == '0') return rc;
    fflush(stdout); fflush(stderr);
    _exit(rc);
}
