// main_cuda.cpp -- driver: the reference's main() (src/main.cpp:10-37) with the one extra factory
// branch INTEGRATION.md describes.  method="FVM_TVD_CUDA" -> the B200 path; method="FVM_TVD" -> the
// reference CPU method (with Cell::flag zeroed), for A/B runs on identical inputs.
// The implicit / DG methods need hypre and stay in the reference's own binary.
#include "fvm_tvd_cuda.h"
#include "tinyxml.h"
#include <ctime>
#include <unistd.h>
#include <string>

int main(int argc, char** argv)
{
	Parallel::init(&argc, &argv);
	if (Parallel::procCount > 1 && Parallel::procId > 0) {
		// every rank runs the reference's serial init (global mesh, regions, save(0)): ranks > 0 do it in a
		// private directory holding links to the inputs, so only rank 0 writes res_*.vtk / task.log in place
		char tmpl[] = "/tmp/cfd2d_rankXXXXXX";
		char * d = mkdtemp(tmpl);
		char cwd[4096];
		if (d && getcwd(cwd, sizeof cwd)) {
			std::string cmd = std::string("for f in '") + cwd + "'/*; do ln -s \"$f\" '" + d + "'/ 2>/dev/null; done";
			if (system(cmd.c_str()) == 0 && chdir(d) == 0 && !getenv("CFD2D_JOB_DIR")) setenv("CFD2D_JOB_DIR", cwd, 1);
		}
	}
	hLog = fopen("task.log", "w");
	const char * xml = argc > 1 ? argv[1] : "task.xml";
	TiXmlDocument doc(xml);
	if (!doc.LoadFile(TIXML_ENCODING_UTF8)) { log("ERROR: %s\n", doc.ErrorDesc()); return doc.ErrorId(); }
	const char * name = doc.FirstChild("task")->ToElement()->Attribute("method");
	Method * m = NULL;
	if (strcmp("FVM_TVD_CUDA", name) == 0) m = new FVM_TVD_CUDA();
	else if (strcmp("FVM_TVD", name) == 0) m = new FVM_TVD_REF0();
	else { log("ERROR: unsupported method '%s' in this driver.\n", name); EXIT(1); }
	m->init((char*)xml);
	struct timespec a, b;
	clock_gettime(CLOCK_MONOTONIC, &a);
	m->run();
	clock_gettime(CLOCK_MONOTONIC, &b);
	log("run() wall time: %.6f s\n", (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec));
	m->done();
	Parallel::done();
	fclose(hLog);
	return 0;
}
