// main_cuda.cpp -- driver: the reference's main() (src/main.cpp:10-37) with the one extra factory
// branch INTEGRATION.md describes.  method="FVM_TVD_CUDA" -> the B200 path; method="FVM_TVD" -> the
// reference CPU method (with Cell::flag zeroed), for A/B runs on identical inputs.
// The implicit / DG methods need hypre and stay in the reference's own binary.
#include "fvm_tvd_cuda.h"
#include "tinyxml.h"
#include <ctime>

int main(int argc, char** argv)
{
	Parallel::init(&argc, &argv);
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
