/* the `sdust` executable: everything lives in liblqcov.so */
#include "lqcov.h"
int main(int argc, char **argv) { return lqcov_sdust_main(argc, argv); }
