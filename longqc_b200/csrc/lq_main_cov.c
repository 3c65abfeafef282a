/* the `minimap2-coverage` executable: everything lives in liblqcov.so */
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include "lqcov.h"
/* The contract of the executable is its stdout bytes and its exit status (lq_exec.py:13-38).  Once both are settled the process
 * leaves through _exit: tearing the CUDA context down in order costs 0.2-0.6 s per process on a B200 box
 * (profiles/r02_cuda_process_costs.log) and gives the caller nothing.  LQCOV_FAST_EXIT=0 takes the orderly way out. */
int main(int argc, char **argv)
{
    setenv("LQCOV_FAST_EXIT", "1", 0);                  /* tells the library not to tear its contexts down either */
    const int rc = lqcov_main(argc, argv);
    const char *fe = getenv("LQCOV_FAST_EXIT");
    if (fe && fe[0] == '0') return rc;
    fflush(stdout); fflush(stderr);
    _exit(rc);
}
