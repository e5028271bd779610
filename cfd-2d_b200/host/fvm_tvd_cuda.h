// fvm_tvd_cuda.h -- the drop-in: a Method (reference src/methods/method.h:6-12) whose hot loop runs
// on a B200 through the C-ABI of include/cfd2d_fvm.h.
//
// It derives from the reference's own FVM_TVD so that EVERYTHING outside the hot loop stays the
// reference's code, unchanged: task.xml parsing, materials/regions, CFDBoundary::create, the mesh
// readers, the edge->boundary binding, the initial state (FVM_TVD::init, fvm_tvd.cpp:7-213) and the
// VTK writer (FVM_TVD::save, :501-600).  Only run() -- the while loop of fvm_tvd.cpp:303-462 -- is
// replaced.
//
// FVM_TVD declares its data members `private:` (fvm_tvd.h:66); a subclass needs them `protected:`.
// A maintainer changes that one word; this out-of-tree build gets the same effect
// without touching the reference by re-labelling the access specifier for this one include
// (access specifiers do not change object layout).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <set>
#define private protected
#include "fvm_tvd.h"
#undef private
#include "../../include/cfd2d_fvm.h"

class FVM_TVD_CUDA : public FVM_TVD
{
public:
	FVM_TVD_CUDA() : h(NULL), device(0), flux(CFD2D_FLUX_GODUNOV), order(2), rank(0), nranks(1) {}
	virtual void init(char * xmlFileName);
	virtual void run();
	virtual void done();
protected:
	void upload();                 // Grid + tables -> cfd2d_fvm_create / set_state / calc_time_step
	void download();               // device state -> ro, ru, rv, re, cTau, Cell::flag
	void fail(const char * what);  // log("ERROR...") + EXIT(1), the reference's error style
	void collectSnapshot(int saveStep);   // cfd2d_fvm_snapshot_end + the reference's save()
	// ---- multi-rank (one process per GPU): Decomp's partition and renumbering, recomputed from the global
	// mesh every rank has read (reference src/methods/decomp.cpp:86-292), checked against mesh/mesh.NNNN.proc
	// when the DECOMP method has written them
	void decomposeMesh();
	bool checkProcFile();
	cfd2d_fvm * h;
	int device, flux, order;
	int rank, nranks;
	std::vector<int> part;                      // METIS part[] of the global mesh (decomp.cpp:104)
	std::vector< std::vector<int> > owned;      // per rank: owned global cells, ascending
	std::vector<int> gCells, gEdges;            // this rank: local -> global cell (owned + halo) / edge
	std::vector<int> recvCount, sendCount, sendInd;   // Grid::recvCount / sendInd of this rank (grid.h:94-96)
};

// The reference CPU method with the one fix-up SURVEY.md F11 documents (Cell::flag is never
// initialised by the UNV reader): used by the driver for A/B runs on the same inputs.
class FVM_TVD_REF0 : public FVM_TVD
{
public:
	virtual void init(char * xmlFileName)
	{
		FVM_TVD::init(xmlFileName);
		for (int i = 0; i < grid.cCount; i++) grid.cells[i].flag = 0;
	}
};
