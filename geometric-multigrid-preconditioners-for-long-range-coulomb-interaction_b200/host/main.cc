// `main <file.prm>` of the reference (src/main.cc:6-121) for the B200 path.
#include <iostream>

#include "step_50.h"

int main(int argc, char *argv[]) { return step50_main(argc, argv, std::cout); }
