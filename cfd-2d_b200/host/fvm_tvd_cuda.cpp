// fvm_tvd_cuda.cpp -- see fvm_tvd_cuda.h.  Host glue only: no numerics here.
#include "fvm_tvd_cuda.h"
#include "tinyxml.h"
#include <dlfcn.h>
#include <unistd.h>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <cstdarg>
#include <ctime>

// The reference's log() (global.cpp:32-46) walks its va_list twice (vprintf, then vfprintf): with arguments
// that is undefined behaviour and crashes on %s.  The glue therefore formats its own messages and passes
// log() a plain string (every '%' doubled), which both passes print verbatim.
static void glueLog(const char * fmt, ...)
{
	char buf[1024], esc[2048];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	size_t k = 0;
	for (size_t i = 0; buf[i] && k + 2 < sizeof esc; i++) { if (buf[i] == '%') esc[k++] = '%'; esc[k++] = buf[i]; }
	esc[k] = 0;
	log(esc);                                                   // root rank only (global.cpp:34)
	if (!Parallel::isRoot() && strncmp(buf, "ERROR", 5) == 0) fprintf(stderr, "[rank %d] %s", Parallel::procId, buf);
}

void FVM_TVD_CUDA::fail(const char * what)
{
	glueLog("ERROR (FVM_TVD_CUDA, %s): %s\n", what, cfd2d_fvm_last_error(h));
	EXIT(1);
}

void FVM_TVD_CUDA::init(char * xmlFileName)
{
	FVM_TVD::init(xmlFileName);            // reference init, incl. calcTimeStep() and save(0)
	for (int i = 0; i < grid.cCount; i++) grid.cells[i].flag = 0;

	// optional element the reference ignores: <gpu device="0" flux="GODUNOV|LAX" order="1|2"/>
	TiXmlDocument doc(xmlFileName);
	if (doc.LoadFile(TIXML_ENCODING_UTF8)) {
		TiXmlNode * task = doc.FirstChild("task");
		TiXmlNode * g = task ? task->FirstChild("gpu") : NULL;
		if (g) {
			TiXmlElement * e = g->ToElement();
			e->Attribute("device", &device);
			e->Attribute("order", &order);
			const char * f = e->Attribute("flux");
			if (f && strcmp(f, "LAX") == 0) flux = CFD2D_FLUX_LAX;
		}
	}
	rank = Parallel::procId;
	nranks = Parallel::procCount > 0 ? Parallel::procCount : 1;
	if (nranks > 1) {
		const char * lr = getenv("LOCAL_RANK");
		device = lr ? atoi(lr) : rank;          // one process per GPU
		decomposeMesh();
	}
	upload();
}

// ---- multi-rank: partition + owned/halo renumbering -------------------------------------------------
// Every rank has read the GLOBAL mesh (FVM_TVD::init), so each recomputes Decomp's maps itself
// (src/methods/decomp.cpp:86-292: dual graph from Cell::neigh -> METIS_PartGraphRecursive with default
// options -> owned cells ascending, halo = non-owned face neighbours in discovery order stably sorted by
// owner, send lists in the receiver's halo order) and builds its share from the global doubles: the
// .proc files print coordinates with %25.15e (not round-trip exact) and their reader re-orients edges per
// rank (grid.cpp:362-411), which would break bitwise equality with the serial run.
typedef int (*metis_recursive_fn)(int*, int*, int*, int*, int*, int*, int*, int*, float*, float*, int*, int*, int*);

void FVM_TVD_CUDA::decomposeMesh()
{
	const int nc = grid.cCount;
	std::vector<int> xadj(nc + 1, 0), adjncy;
	adjncy.reserve(3 * (size_t)nc);
	for (int i = 0; i < nc; i++) {                               // decomp.cpp:86-102
		for (int k = 0; k < 3; k++) if (grid.cells[i].neigh[k] >= 0) adjncy.push_back(grid.cells[i].neigh[k]);
		xadj[i + 1] = (int)adjncy.size();
	}
	part.assign(nc, 0);
	{
		std::string lib = getenv("CFD2D_METIS_LIB") ? getenv("CFD2D_METIS_LIB") : "";
		if (lib.empty()) {
			char exe[4096]; ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 1);
			std::string d = n > 0 ? std::string(exe, (size_t)n) : std::string(".");
			d = d.substr(0, d.find_last_of('/'));                  // .../host/_build
			lib = d + "/../../third_party/_metis/libmetis.so";
		}
		void * hm = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
		if (!hm) { glueLog("ERROR (FVM_TVD_CUDA): cannot load the bundled METIS (%s): %s\n", lib.c_str(), dlerror()); EXIT(1); }
		metis_recursive_fn f = (metis_recursive_fn)dlsym(hm, "METIS_PartGraphRecursive");
		if (!f) { glueLog("ERROR (FVM_TVD_CUDA): METIS_PartGraphRecursive not found in %s\n", lib.c_str()); EXIT(1); }
		int n = nc, ncon = 1, np = nranks, objval = 0;
		int rc = f(&n, &ncon, xadj.data(), adjncy.data(), NULL, NULL, NULL, &np, NULL, NULL, NULL, &objval, part.data());   // decomp.cpp:104
		if (rc != 1) { glueLog("ERROR (FVM_TVD_CUDA): METIS_PartGraphRecursive returned %d\n", rc); EXIT(1); }
	}
	owned.assign(nranks, std::vector<int>());
	for (int i = 0; i < nc; i++) owned[part[i]].push_back(i);   // ascending global id, decomp.cpp:165-171
	// halo of every rank, in the order Decomp builds it (cell ascending, neigh slot 0..2, duplicates kept;
	// then its bubble sort by owner == a stable sort), decomp.cpp:172-208
	std::vector< std::vector<int> > halo(nranks);
	for (int p = 0; p < nranks; p++) {
		std::vector<int> & hp = halo[p];
		for (size_t q = 0; q < owned[p].size(); q++) {
			const Cell & c = grid.cells[owned[p][q]];
			for (int k = 0; k < 3; k++) if (c.neigh[k] >= 0 && part[c.neigh[k]] != p) hp.push_back(c.neigh[k]);
		}
		std::stable_sort(hp.begin(), hp.end(), [&](int a, int b) { return part[a] < part[b]; });
	}
	const int ncLoc = (int)owned[rank].size();
	gCells = owned[rank];
	gCells.insert(gCells.end(), halo[rank].begin(), halo[rank].end());
	std::vector<int> lCells(nc, -1);
	for (int i = 0; i < (int)gCells.size(); i++) lCells[gCells[i]] = i;          // last occurrence wins, decomp.cpp:209-211
	recvCount.assign(nranks, 0);
	for (size_t i = 0; i < halo[rank].size(); i++) recvCount[part[halo[rank][i]]]++;
	// owner-side send lists in the receiver's halo order, decomp.cpp:284-292
	sendCount.assign(nranks, 0);
	sendInd.clear();
	for (int p = 0; p < nranks; p++) {
		if (p == rank) continue;
		for (size_t i = 0; i < halo[p].size(); i++)
			if (part[halo[p][i]] == rank) { sendInd.push_back(lCells[halo[p][i]]); sendCount[p]++; }
	}
	// edges touching an owned cell, ascending global id, decomp.cpp:217-236
	std::vector<char> touch(grid.eCount, 0);
	for (int i = 0; i < ncLoc; i++) for (int k = 0; k < 3; k++) touch[grid.cells[gCells[i]].edgesInd[k]] = 1;
	gEdges.clear();
	for (int e = 0; e < grid.eCount; e++) if (touch[e]) gEdges.push_back(e);
	glueLog("FVM_TVD_CUDA rank %d/%d: %d owned + %d halo cells, %d edges\n", rank, nranks, ncLoc, (int)gCells.size() - ncLoc, (int)gEdges.size());
	if (!checkProcFile()) EXIT(1);
}

// When the DECOMP method has been run (mesh/mesh.NNNN.proc, format decomp.cpp:295-334), its integer maps
// must be the ones computed above: counts, Grid::recvCount and Grid::sendInd (grid.cpp:413-443).
bool FVM_TVD_CUDA::checkProcFile()
{
	char name[64];
	sprintf(name, "mesh/mesh.%04d.proc", rank);
	std::ifstream f(name);
	if (!f) return true;                                       // not decomposed on disk: nothing to compare
	std::string line;
	auto next_line = [&](std::string & out) { while (std::getline(f, out)) { if (out.find_first_not_of(" \t\r") != std::string::npos) return true; } return false; };
	long a = 0, b = 0;
	bool ok = true;
	auto counts = [&](long & x, long & y) { if (!next_line(line)) return false; std::istringstream is(line); return bool(is >> x >> y); };
	auto skip = [&](long n) { for (long i = 0; i < n; i++) if (!next_line(line)) return false; return true; };
	if (!counts(a, b) || !skip(b)) ok = false;                 // nodes
	long cC = 0, cCEx = 0;
	if (ok && (!counts(cC, cCEx) || !skip(cCEx))) ok = false;  // cells
	long eC = 0, eCEx = 0;
	if (ok && (!counts(eC, eCEx) || !skip(eCEx))) ok = false;  // edges
	if (!ok) { glueLog("ERROR (FVM_TVD_CUDA): cannot parse %s\n", name); return false; }
	const long ncLoc = (long)owned[rank].size();
	if (cC != ncLoc || cCEx != (long)gCells.size() || eC != (long)gEdges.size()) {
		glueLog("ERROR (FVM_TVD_CUDA): %s has cCount=%ld cCountEx=%ld eCount=%ld, this partition %ld %ld %ld\n", name, cC, cCEx, eC,
		    ncLoc, (long)gCells.size(), (long)gEdges.size());
		return false;
	}
	if (!next_line(line)) return false;
	{
		std::istringstream is(line);
		for (int p = 0; p < nranks; p++) { long v = -1; is >> v; if (v != recvCount[p]) { glueLog("ERROR (FVM_TVD_CUDA): %s recvCount[%d]=%ld != %d\n", name, p, v, recvCount[p]); return false; } }
	}
	size_t off = 0;
	for (int p = 0; p < nranks; p++) {
		if (!next_line(line)) break;
		std::istringstream is(line);
		long pp = -1, n = -1;
		is >> pp >> n;
		if (pp != p || n != (p == rank ? 0 : sendCount[p])) { glueLog("ERROR (FVM_TVD_CUDA): %s send list of rank %d has %ld entries, expected %d\n", name, p, n, p == rank ? 0 : sendCount[p]); return false; }
		for (long i = 0; i < n; i++) {
			long v = -1;
			if (!(is >> v)) { if (!next_line(line)) return false; is.clear(); is.str(line); is >> v; }
			if (v != sendInd[off + i]) { glueLog("ERROR (FVM_TVD_CUDA): %s sendInd[%d][%ld]=%ld != %d\n", name, p, i, v, sendInd[off + i]); return false; }
		}
		off += (size_t)(n > 0 ? n : 0);
	}
	glueLog("FVM_TVD_CUDA rank %d: partition maps equal %s (cCount, cCountEx, eCount, recvCount, sendInd)\n", rank, name);
	return true;
}

void FVM_TVD_CUDA::upload()
{
	const bool multi = nranks > 1;
	const int ncOwn = multi ? (int)owned[rank].size() : grid.cCount;
	const int nc = multi ? (int)gCells.size() : grid.cCount, ne = multi ? (int)gEdges.size() : grid.eCount;
	std::vector<int> lCells, lEdges;
	if (multi) {
		lCells.assign(grid.cCount, -1); lEdges.assign(grid.eCount, -1);
		for (int i = 0; i < nc; i++) lCells[gCells[i]] = i;
		for (int i = 0; i < ne; i++) lEdges[gEdges[i]] = i;
	}
	auto GC = [&](int i) { return multi ? gCells[i] : i; };
	auto GE = [&](int i) { return multi ? gEdges[i] : i; };
	std::vector<double> cS(nc), cx(nc), cy(nc), enx(ne), eny(ne), el(ne), egp(4 * (size_t)ne);
	std::vector<int> cmat(nc), cedges(3 * (size_t)nc), ec1(ne), ec2(ne), ebc(ne);
	for (int i = 0; i < nc; i++) {
		Cell & c = grid.cells[GC(i)];
		cS[i] = c.S; cx[i] = c.c.x; cy[i] = c.c.y;
		cmat[i] = getRegion(c.typeName).matId;                     // getMaterial(), fvm_tvd.cpp:788-792
		if (i < ncOwn) for (int k = 0; k < 3; k++) cedges[3 * (size_t)i + k] = multi ? lEdges[c.edgesInd[k]] : c.edgesInd[k];
	}
	for (int i = 0; i < ne; i++) {
		Edge & e = grid.edges[GE(i)];
		// every rank keeps the GLOBAL orientation (c1 -> c2, normal, Gauss-point order) of the edge
		ec1[i] = multi ? lCells[e.c1] : e.c1; ec2[i] = (multi && e.c2 >= 0) ? lCells[e.c2] : e.c2;
		enx[i] = e.n.x; eny[i] = e.n.y; el[i] = e.l;
		if (e.cCount != 3) {   // the Triangle reader gives cCount=1 => no Gauss points => no flux
			glueLog("ERROR (FVM_TVD_CUDA): edge %d has %d points; the path needs centre + 2 Gauss points (salome_unv meshes)\n", i, e.cCount);
			EXIT(1);
		}
		egp[4 * (size_t)i + 0] = e.c[1].x; egp[4 * (size_t)i + 1] = e.c[1].y;
		egp[4 * (size_t)i + 2] = e.c[2].x; egp[4 * (size_t)i + 3] = e.c[2].y;
		ebc[i] = -1;
		for (int b = 0; b < bCount; b++) if (boundaries[b] == e.bnd) ebc[i] = b;
	}
	std::vector<double> mM(matCount), mCp(matCount), bpar(4 * (size_t)(bCount ? bCount : 1), 0.0);
	std::vector<int> bkind(bCount ? bCount : 1, 0);
	for (int i = 0; i < matCount; i++) { mM[i] = materials[i].M; mCp[i] = materials[i].Cp; }
	for (int b = 0; b < bCount; b++) {
		CFDBoundary * bc = boundaries[b];
		if (dynamic_cast<CFDBndInlet*>(bc)) bkind[b] = CFD2D_BC_INLET;
		else if (dynamic_cast<CFDBndOutlet*>(bc)) bkind[b] = CFD2D_BC_OUTLET;
		else bkind[b] = CFD2D_BC_WALL;                              // slip and "no-slip", bnd_cond.cpp:50-62
		for (int k = 0; k < 4 && k < bc->parCount; k++) bpar[4 * (size_t)b + k] = bc->par[k];
	}
	cfd2d_mesh m;
	m.nc = ncOwn; m.nc_ex = nc; m.ne = ne;
	m.cell_S = cS.data(); m.cell_cx = cx.data(); m.cell_cy = cy.data(); m.cell_mat = cmat.data();
	m.cell_edges = cedges.data(); m.edge_c1 = ec1.data(); m.edge_c2 = ec2.data();
	m.edge_nx = enx.data(); m.edge_ny = eny.data(); m.edge_l = el.data(); m.edge_gp = egp.data(); m.edge_bc = ebc.data();
	cfd2d_phys p;
	p.nmat = matCount; p.mat_M = mM.data(); p.mat_Cp = mCp.data();
	p.nbc = bCount; p.bc_kind = bkind.data(); p.bc_par = bpar.data();
	p.limits[0] = limitRmin; p.limits[1] = limitRmax; p.limits[2] = limitPmin; p.limits[3] = limitPmax; p.limits[4] = limitUmax;
	cfd2d_ctrl c;
	c.CFL = CFL; c.TAU = TAU; c.steady = STEADY ? 1 : 0; c.flux = flux; c.order = order; c.max_newton = 0;
	cfd2d_halo halo;
	char ncclId[128];
	if (multi) {
		// the 128-byte ncclUniqueId: created by rank 0, handed to the others through a file in the job directory
		// (a site with MPI would MPI_Bcast it through Parallel; this build has no MPI)
		std::string dir = getenv("CFD2D_JOB_DIR") ? getenv("CFD2D_JOB_DIR") : ".";
		std::string job = getenv("CFD2D_JOB_ID") ? getenv("CFD2D_JOB_ID") : "0";
		std::string path = dir + "/.cfd2d_nccl_id." + job, tmp = path + ".tmp";
		if (rank == 0) {
			if (cfd2d_nccl_get_unique_id(ncclId) != 0) { glueLog("ERROR (FVM_TVD_CUDA): ncclGetUniqueId failed\n"); EXIT(1); }
			FILE * f = fopen(tmp.c_str(), "wb");
			if (!f || fwrite(ncclId, 1, 128, f) != 128) { glueLog("ERROR (FVM_TVD_CUDA): cannot write %s\n", tmp.c_str()); EXIT(1); }
			fclose(f);
			rename(tmp.c_str(), path.c_str());
		} else {
			int tries = 0;
			for (;; tries++) {
				FILE * f = fopen(path.c_str(), "rb");
				if (f) { size_t n = fread(ncclId, 1, 128, f); fclose(f); if (n == 128) break; }
				if (tries > 1200) { glueLog("ERROR (FVM_TVD_CUDA): rank %d: no NCCL id at %s after 120 s\n", rank, path.c_str()); EXIT(1); }
				usleep(100000);
			}
		}
		halo.rank = rank; halo.nranks = nranks;
		halo.recv_count = recvCount.data(); halo.send_count = sendCount.data();
		halo.send_ind = sendInd.empty() ? NULL : sendInd.data();
		halo.nccl_unique_id = ncclId;
		halo.cell_gid = gCells.data();
		if (sendInd.empty()) { static int zero = 0; halo.send_ind = &zero; }
	}
	int rc = cfd2d_fvm_create(&m, &p, &c, multi ? &halo : NULL, device, &h);
	if (rc != 0) { h = NULL; fail("create"); }
	std::vector<uint32_t> fl(ncOwn, 0u);
	if (multi) {
		std::vector<double> a(ncOwn), b(ncOwn), cc(ncOwn), d(ncOwn);
		for (int i = 0; i < ncOwn; i++) { a[i] = ro[gCells[i]]; b[i] = ru[gCells[i]]; cc[i] = rv[gCells[i]]; d[i] = re[gCells[i]]; }
		if (cfd2d_fvm_set_state(h, a.data(), b.data(), cc.data(), d.data(), fl.data()) != 0) fail("set_state");
	} else
	if (cfd2d_fvm_set_state(h, ro, ru, rv, re, fl.data()) != 0) fail("set_state");
	double tau = 0.0;
	if (cfd2d_fvm_calc_time_step(h, &tau) != 0) fail("calc_time_step");   // == TAU from the CPU calcTimeStep
	TAU = tau;
	glueLog("FVM_TVD_CUDA: %d cells, %d edges on device %d (%s)\n", nc, ne, device, cfd2d_version());
	if (multi && rank == 0) { std::string dir = getenv("CFD2D_JOB_DIR") ? getenv("CFD2D_JOB_DIR") : "."; std::string job = getenv("CFD2D_JOB_ID") ? getenv("CFD2D_JOB_ID") : "0"; remove((dir + "/.cfd2d_nccl_id." + job).c_str()); }
}

void FVM_TVD_CUDA::download()
{
	if (nranks > 1) {
		// the owned cells of every rank -> rank 0 (NCCL, cfd2d_fvm_gather_state), scattered into the global arrays
		// the reference's writer reads; the other ranks only refresh their own cells
		std::vector<int> counts(nranks);
		size_t total = 0;
		for (int p = 0; p < nranks; p++) { counts[p] = (int)owned[p].size(); total += owned[p].size(); }
		std::vector<double> a, b, c, d, t;
		std::vector<uint32_t> fl;
		if (rank == 0) { a.resize(total); b.resize(total); c.resize(total); d.resize(total); t.resize(total); fl.resize(total); }
		if (cfd2d_fvm_gather_state(h, 0, counts.data(), rank == 0 ? a.data() : NULL, rank == 0 ? b.data() : NULL, rank == 0 ? c.data() : NULL,
		                           rank == 0 ? d.data() : NULL, rank == 0 ? t.data() : NULL, rank == 0 ? fl.data() : NULL) != 0) fail("gather_state");
		if (rank == 0) {
			size_t k = 0;
			for (int p = 0; p < nranks; p++)
				for (size_t i = 0; i < owned[p].size(); i++, k++) {
					const int g = owned[p][i];
					ro[g] = a[k]; ru[g] = b[k]; rv[g] = c[k]; re[g] = d[k]; cTau[g] = t[k]; grid.cells[g].flag = fl[k];
				}
		}
		return;
	}
	std::vector<uint32_t> fl(grid.cCount);
	if (cfd2d_fvm_get_state(h, ro, ru, rv, re, cTau, fl.data()) != 0) fail("get_state");
	for (int i = 0; i < grid.cCount; i++) grid.cells[i].flag = fl[i];
}

// The reference loop (fvm_tvd.cpp:303-462) with the device doing the steps.  Output is pipelined:
// at a save step the state is snapshotted on the device (cfd2d_fvm_snapshot_begin: unpack + D2H on a
// copy stream), the NEXT chunk of steps is enqueued, and only then the reference's own VTK writer
// (save(), :501-600 -- unchanged, so res_*.vtk stays byte-identical) runs on the host, under the GPU.
void FVM_TVD_CUDA::collectSnapshot(int saveStep)
{
	std::vector<uint32_t> fl(grid.cCount);
	if (cfd2d_fvm_snapshot_end(h, ro, ru, rv, re, cTau, fl.data()) != 0) fail("snapshot_end");
	for (int i = 0; i < grid.cCount; i++) grid.cells[i].flag = fl[i];
	save(saveStep);                          // the reference's own VTK writer
}

void FVM_TVD_CUDA::run()
{
	double       t    = 0.0;
	unsigned int step = 0;
	int pendingSave = -1;
	struct timespec t0;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	while (t < TMAX && step < (unsigned int)STEP_MAX)
	{
		// how many steps until the next save / log line / end, exactly as the reference loop counts them
		unsigned int n = (unsigned int)STEP_MAX - step;
		unsigned int toSave = FILE_SAVE_STEP - step % FILE_SAVE_STEP;
		unsigned int toLog  = PRINT_STEP - step % PRINT_STEP;
		if (toSave < n) n = toSave;
		if (toLog < n) n = toLog;
		if (!STEADY) {                       // t += TAU per step (fvm_tvd.cpp:313): same additions, counted ahead
			double tt = t; unsigned int k = 0;
			while (k < n && tt < TMAX) { tt += TAU; k++; }
			n = k;
			t = tt;
		}
		if (cfd2d_fvm_step_async(h, (int)n) != 0) fail("step");        // the GPU starts the chunk ...
		if (pendingSave >= 0) { collectSnapshot(pendingSave); pendingSave = -1; }   // ... the file is written under it
		if (cfd2d_fvm_sync(h) != 0) fail("step");
		step += n;
		if (step % FILE_SAVE_STEP == 0)
		{
			if (nranks > 1) {                // gather over NCCL, rank 0 writes the (global) file
				download();
				if (rank == 0) save(step);
			} else {
				if (cfd2d_fvm_snapshot_begin(h) != 0) fail("snapshot_begin");
				pendingSave = (int)step;
			}
		}
		if (step % PRINT_STEP == 0)
		{
			log("step: %d\t\ttime step: %.16f\n", step, t);
		}
	}
	if (pendingSave >= 0) collectSnapshot(pendingSave);
	download();
	{
		// throughput in the project's metric: cell-updates/s per RK stage, and the 320 B/cell-stage
		// algorithmic bandwidth it corresponds to; wall clock of the whole loop, output included
		struct timespec t1;
		clock_gettime(CLOCK_MONOTONIC, &t1);
		const double wall = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
		const double cu = 2.0 * (double)grid.cCount * (double)step / (wall > 0 ? wall : 1);
		glueLog("FVM_TVD_CUDA: %u steps of %d cells on %d GPU(s) in %.6f s (output included): %.4g cell-updates/s, %.1f GB/s algorithmic (320 B/cell-stage)\n",
		        step, grid.cCount, nranks, wall, cu, cu * 320.0 / 1e9);
	}
}

void FVM_TVD_CUDA::done()
{
	cfd2d_fvm_destroy(h);
	h = NULL;
	FVM_TVD::done();
}
