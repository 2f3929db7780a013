#include "check.h"
int main(int argc, char **argv) { return check::run(argc, argv); }
