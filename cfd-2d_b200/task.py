"""task.xml for the FVM_TVD method: the schema ``FVM_TVD::init`` consumes
(reference ``src/methods/fvm_tvd.cpp:24-128``, boundary part ``src/bnd_cond.cpp:10-72``;
SURVEY.md Appendix B), as dataclasses with a writer and a reader.

The reader mirrors the reference's attribute-by-name lookups: every scalar lives in a
``value="..."`` attribute, ``<material id>`` inside a region is an ARRAY INDEX into the materials
list (``fvm_tvd.cpp:90``), regions bind to cells by group name, boundaries bind to edges by group
name.  Unknown boundary types raise, like ``CFDBoundary::create`` (``bnd_cond.cpp:64-66``).
"""
from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET

GR = 8.314472  # Material::gR, src/global.cpp:6

BOUND_INLET = "BOUND_INLET"
BOUND_OUTLET = "BOUND_OUTLET"
BOUND_WALL_SLIP = "BOUND_WALL_SLIP"
BOUND_WALL_NO_SLIP = "BOUND_WALL_NO_SLIP"   # instantiates the slip class (bnd_cond.cpp:50-55)

# kinds in the C-ABI BC table (include/cfd2d_fvm.h)
BC_INLET, BC_OUTLET, BC_WALL = 1, 2, 3
_KIND = {BOUND_INLET: BC_INLET, BOUND_OUTLET: BC_OUTLET, BOUND_WALL_SLIP: BC_WALL, BOUND_WALL_NO_SLIP: BC_WALL}


@dataclasses.dataclass
class Material:
    name: str = "air"
    M: float = 0.02898
    Cp: float = 1004.5
    K: float = 0.0
    ML: float = 0.0

    @property
    def Cv(self) -> float:
        return self.Cp - GR / self.M        # Material::URS, global.cpp:11

    @property
    def gamma(self) -> float:
        return self.Cp / self.Cv            # global.cpp:12


@dataclasses.dataclass
class Region:
    name: str
    mat_id: int = 0
    Vx: float = 0.0
    Vy: float = 0.0
    T: float = 300.0
    P: float = 1.0e5
    cell_type: int = 0


@dataclasses.dataclass
class BoundCond:
    name: str
    type: str = BOUND_WALL_SLIP
    Vx: float = 0.0
    Vy: float = 0.0
    T: float = 0.0
    P: float = 0.0
    edge_type: int = 0

    @property
    def kind(self) -> int:
        if self.type not in _KIND:
            raise ValueError(f"Unknown boundary type '{self.type}' specified.")
        return _KIND[self.type]

    @property
    def par(self):
        return [self.Vx, self.Vy, self.T, self.P] if self.kind == BC_INLET else [0.0, 0.0, 0.0, 0.0]


@dataclasses.dataclass
class Task:
    steady: int = 0
    TAU: float = 1.0e-6
    TMAX: float = 1.0e10
    CFL: float = 0.3
    STEP_MAX: int = 100
    FILE_OUTPUT_STEP: int = 1000000000
    LOG_OUTPUT_STEP: int = 1000000000
    ro_min: float = 1.0e-6
    ro_max: float = 1.0e6
    p_min: float = 1.0e-3
    p_max: float = 1.0e12
    u_max: float = 1.0e6
    materials: list = dataclasses.field(default_factory=lambda: [Material()])
    regions: list = dataclasses.field(default_factory=list)
    boundaries: list = dataclasses.field(default_factory=list)
    mesh_name: str = "mesh.unv"
    mesh_type: str = "salome_unv"
    method: str = "FVM_TVD"

    @property
    def limits(self):
        return [self.ro_min, self.ro_max, self.p_min, self.p_max, self.u_max]


def _v(x) -> str:
    return repr(float(x))


def write_task_xml(path: str, t: Task) -> None:
    root = ET.Element("task", method=t.method)
    c = ET.SubElement(root, "control")
    ET.SubElement(c, "STEADY", value=str(int(t.steady)))
    ET.SubElement(c, "TAU", value=_v(t.TAU))
    ET.SubElement(c, "TMAX", value=_v(t.TMAX))
    ET.SubElement(c, "CFL", value=_v(t.CFL))
    ET.SubElement(c, "STEP_MAX", value=str(int(t.STEP_MAX)))
    ET.SubElement(c, "FILE_OUTPUT_STEP", value=str(int(t.FILE_OUTPUT_STEP)))
    ET.SubElement(c, "LOG_OUTPUT_STEP", value=str(int(t.LOG_OUTPUT_STEP)))
    lim = ET.SubElement(root, "limits")
    ET.SubElement(lim, "ro", min=_v(t.ro_min), max=_v(t.ro_max))
    ET.SubElement(lim, "p", min=_v(t.p_min), max=_v(t.p_max))
    ET.SubElement(lim, "u", max=_v(t.u_max))
    mats = ET.SubElement(root, "materials", count=str(len(t.materials)))
    for i, m in enumerate(t.materials):
        me = ET.SubElement(mats, "material", id=str(i))
        ET.SubElement(me, "name").text = m.name
        p = ET.SubElement(me, "parameters")
        ET.SubElement(p, "M", value=_v(m.M))
        ET.SubElement(p, "Cp", value=_v(m.Cp))
        ET.SubElement(p, "K", value=_v(m.K))
        ET.SubElement(p, "ML", value=_v(m.ML))
    regs = ET.SubElement(root, "regions", count=str(len(t.regions)))
    for i, r in enumerate(t.regions):
        re_ = ET.SubElement(regs, "region", id=str(i))
        ET.SubElement(re_, "material", id=str(r.mat_id))
        ET.SubElement(re_, "cell", type=str(r.cell_type))
        ET.SubElement(re_, "name").text = r.name
        p = ET.SubElement(re_, "parameters")
        ET.SubElement(p, "Vx", value=_v(r.Vx))
        ET.SubElement(p, "Vy", value=_v(r.Vy))
        ET.SubElement(p, "T", value=_v(r.T))
        ET.SubElement(p, "P", value=_v(r.P))
    bs = ET.SubElement(root, "boundaries")
    for b in t.boundaries:
        be = ET.SubElement(bs, "boundCond", edgeType=str(b.edge_type))
        ET.SubElement(be, "name").text = b.name
        ET.SubElement(be, "type").text = b.type
        p = ET.SubElement(be, "parameters")
        if b.type == BOUND_INLET:
            ET.SubElement(p, "Vx", value=_v(b.Vx))
            ET.SubElement(p, "Vy", value=_v(b.Vy))
            ET.SubElement(p, "T", value=_v(b.T))
            ET.SubElement(p, "P", value=_v(b.P))
    me = ET.SubElement(root, "mesh")
    ET.SubElement(me, "name", value=t.mesh_name)
    ET.SubElement(me, "filesType", value=t.mesh_type)
    ET.indent(root)
    ET.ElementTree(root).write(path, encoding="utf-8", xml_declaration=True)


def _attr(node, name, conv=float):
    return conv(node.attrib[name])


def read_task_xml(path: str) -> Task:
    """Parse task.xml the way FVM_TVD::init does (fvm_tvd.cpp:24-128)."""
    root = ET.parse(path).getroot()
    if root.tag != "task":
        raise ValueError("task.xml: root element must be <task>")
    t = Task(method=root.attrib.get("method", "FVM_TVD"))
    c = root.find("control")
    t.steady = 0 if int(c.find("STEADY").attrib["value"]) == 0 else 1
    t.TAU = _attr(c.find("TAU"), "value")
    t.TMAX = _attr(c.find("TMAX"), "value")
    t.CFL = _attr(c.find("CFL"), "value")
    t.STEP_MAX = _attr(c.find("STEP_MAX"), "value", int)
    t.FILE_OUTPUT_STEP = _attr(c.find("FILE_OUTPUT_STEP"), "value", int)
    t.LOG_OUTPUT_STEP = _attr(c.find("LOG_OUTPUT_STEP"), "value", int)
    lim = root.find("limits")
    t.ro_min = _attr(lim.find("ro"), "min")
    t.ro_max = _attr(lim.find("ro"), "max")
    t.p_min = _attr(lim.find("p"), "min")
    t.p_max = _attr(lim.find("p"), "max")
    t.u_max = _attr(lim.find("u"), "max")
    t.materials = []
    for me in root.find("materials").findall("material"):
        p = me.find("parameters")
        t.materials.append(Material(name=(me.findtext("name") or "").strip(),
                                    M=_attr(p.find("M"), "value"), Cp=_attr(p.find("Cp"), "value"),
                                    K=_attr(p.find("K"), "value"), ML=_attr(p.find("ML"), "value")))
    t.regions = []
    for re_ in root.find("regions").findall("region"):
        p = re_.find("parameters")
        t.regions.append(Region(name=(re_.findtext("name") or "").strip(),
                                mat_id=int(re_.find("material").attrib["id"]),
                                cell_type=int(re_.find("cell").attrib.get("type", 0)),
                                Vx=_attr(p.find("Vx"), "value"), Vy=_attr(p.find("Vy"), "value"),
                                T=_attr(p.find("T"), "value"), P=_attr(p.find("P"), "value")))
    t.boundaries = []
    for be in root.find("boundaries").findall("boundCond"):
        typ = (be.findtext("type") or "").strip()
        b = BoundCond(name=(be.findtext("name") or "").strip(), type=typ,
                      edge_type=int(be.attrib.get("edgeType", 0)))
        _ = b.kind  # raises on unknown types, like CFDBoundary::create
        if typ == BOUND_INLET:
            p = be.find("parameters")
            for k in ("Vx", "Vy", "T", "P"):
                node = p.find(k) if p is not None else None
                if node is None:
                    raise ValueError(f"Parameter '{k}' isn't specified  for BOUND_INLET.")
                setattr(b, k, float(node.attrib["value"]))
        t.boundaries.append(b)
    me = root.find("mesh")
    t.mesh_name = me.find("name").attrib["value"]
    t.mesh_type = me.find("filesType").attrib["value"]
    return t


def region_state(t: Task, r: Region):
    """Conservative state of a region: URS(2) then URS(1) then convertParToCons
    (fvm_tvd.cpp:90-92, :795-801; global.cpp:9-30).  Same operation order as the reference."""
    m = t.materials[r.mat_id]
    Cv = m.Cp - GR / m.M
    gam = m.Cp / Cv
    p, T, u, v = r.P, r.T, r.Vx, r.Vy
    rho = p * m.M / (T * GR)
    e = p / (rho * (gam - 1))
    ro = rho
    ru = rho * u
    rv = rho * v
    re_ = rho * (e + 0.5 * (u * u + v * v))
    return ro, ru, rv, re_
