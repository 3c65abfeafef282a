/* the `minimap2-coverage` executable: everything lives in liblqcov.so */
#include "lqcov.h"
int main(int argc, char **argv) { return lqcov_main(argc, argv); }
