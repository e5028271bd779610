/* ref_decomp_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 * Window on the UNMODIFIED reference mesh splitter (class Decomp, src/methods/decomp.cpp:69-335)
 * linked with the reference's bundled METIS 5.1.0: runs it in <workdir> on <xml> (which must carry
 * <decomp><processors value="k"/></decomp> and the <mesh> element); it writes parts.vtk and
 * mesh/mesh.NNNN.proc there.  tests/test_decomp.py parses those files and compares the integer maps
 * with cfd-2d_b200/decomp.py bit for bit. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>
#include "decomp.h"

extern "C" int ref_decomp_run(const char* workdir, const char* xml) {
    if (chdir(workdir) != 0) return -1;
    if (!hLog) hLog = fopen("task.log", "w");
    Parallel::procCount = 1;
    Parallel::procId = 0;
    fflush(stdout);
    int saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1); close(nul);
    Decomp d;
    d.init((char*)xml);
    d.run();
    fflush(stdout);
    dup2(saved, 1); close(saved);
    return 0;
}
