// TEST INFRASTRUCTURE ONLY: the reference's own habdec::IQSource_File<float> (header-only,
// /root/reference/code/IQSource/IQSource_File.h) driven by oracle/iqsource_script.h.  Built into oracle/_ref/.
#include "IQSource/IQSource_File.h"
#include "iqsource_script.h"
int main(int argc, char** argv) { return argc > 1 ? iqsource_script<habdec::IQSource_File<float>>(argv[1]) : 2; }
