// fvm_tvd_cuda.cpp -- see fvm_tvd_cuda.h.  Host glue only: no numerics here.
#include "fvm_tvd_cuda.h"
#include "tinyxml.h"

void FVM_TVD_CUDA::fail(const char * what)
{
	log("ERROR (FVM_TVD_CUDA, %s): %s\n", what, cfd2d_fvm_last_error(h));
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
	upload();
}

void FVM_TVD_CUDA::upload()
{
	const int nc = grid.cCount, ne = grid.eCount;
	std::vector<double> cS(nc), cx(nc), cy(nc), enx(ne), eny(ne), el(ne), egp(4 * (size_t)ne);
	std::vector<int> cmat(nc), cedges(3 * (size_t)nc), ec1(ne), ec2(ne), ebc(ne);
	for (int i = 0; i < nc; i++) {
		Cell & c = grid.cells[i];
		cS[i] = c.S; cx[i] = c.c.x; cy[i] = c.c.y;
		cmat[i] = getRegion(c.typeName).matId;                     // getMaterial(), fvm_tvd.cpp:788-792
		for (int k = 0; k < 3; k++) cedges[3 * (size_t)i + k] = c.edgesInd[k];
	}
	for (int i = 0; i < ne; i++) {
		Edge & e = grid.edges[i];
		ec1[i] = e.c1; ec2[i] = e.c2;
		enx[i] = e.n.x; eny[i] = e.n.y; el[i] = e.l;
		if (e.cCount != 3) {   // the Triangle reader gives cCount=1 => no Gauss points => no flux
			log("ERROR (FVM_TVD_CUDA): edge %d has %d points; the path needs centre + 2 Gauss points (salome_unv meshes)\n", i, e.cCount);
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
	m.nc = nc; m.nc_ex = nc; m.ne = ne;
	m.cell_S = cS.data(); m.cell_cx = cx.data(); m.cell_cy = cy.data(); m.cell_mat = cmat.data();
	m.cell_edges = cedges.data(); m.edge_c1 = ec1.data(); m.edge_c2 = ec2.data();
	m.edge_nx = enx.data(); m.edge_ny = eny.data(); m.edge_l = el.data(); m.edge_gp = egp.data(); m.edge_bc = ebc.data();
	cfd2d_phys p;
	p.nmat = matCount; p.mat_M = mM.data(); p.mat_Cp = mCp.data();
	p.nbc = bCount; p.bc_kind = bkind.data(); p.bc_par = bpar.data();
	p.limits[0] = limitRmin; p.limits[1] = limitRmax; p.limits[2] = limitPmin; p.limits[3] = limitPmax; p.limits[4] = limitUmax;
	cfd2d_ctrl c;
	c.CFL = CFL; c.TAU = TAU; c.steady = STEADY ? 1 : 0; c.flux = flux; c.order = order; c.max_newton = 0;
	int rc = cfd2d_fvm_create(&m, &p, &c, NULL, device, &h);
	if (rc != 0) { h = NULL; fail("create"); }
	std::vector<uint32_t> fl(nc, 0u);
	if (cfd2d_fvm_set_state(h, ro, ru, rv, re, fl.data()) != 0) fail("set_state");
	double tau = 0.0;
	if (cfd2d_fvm_calc_time_step(h, &tau) != 0) fail("calc_time_step");   // == TAU from the CPU calcTimeStep
	TAU = tau;
	log("FVM_TVD_CUDA: %d cells, %d edges on device %d (%s)\n", nc, ne, device, cfd2d_version());
}

void FVM_TVD_CUDA::download()
{
	std::vector<uint32_t> fl(grid.cCount);
	if (cfd2d_fvm_get_state(h, ro, ru, rv, re, cTau, fl.data()) != 0) fail("get_state");
	for (int i = 0; i < grid.cCount; i++) grid.cells[i].flag = fl[i];
}

void FVM_TVD_CUDA::run()
{
	double       t    = 0.0;
	unsigned int step = 0;
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
		if (cfd2d_fvm_step(h, (int)n) != 0) fail("step");
		step += n;
		if (step % FILE_SAVE_STEP == 0)
		{
			download();
			save(step);                      // the reference's own VTK writer
		}
		if (step % PRINT_STEP == 0)
		{
			log("step: %d\t\ttime step: %.16f\n", step, t);
		}
	}
	download();
}

void FVM_TVD_CUDA::done()
{
	cfd2d_fvm_destroy(h);
	h = NULL;
	FVM_TVD::done();
}
