// habdec_b200::IQSourceFile driven by the same script as the reference's file source (oracle/iqsource_script.h)
#include "habdec_b200/IQSource.hpp"
#include "iqsource_script.h"
int main(int argc, char** argv) { return argc > 1 ? iqsource_script<habdec_b200::IQSourceFile>(argv[1]) : 2; }
