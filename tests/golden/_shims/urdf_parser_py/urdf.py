"""xml.etree-based stand-in for urdf_parser_py.urdf (test infrastructure, see __init__).

Defaults mirror urdf_parser_py: origin xyz/rpy default to [0,0,0]; a joint without <axis>
has axis None... (urdf_parser_py actually defaults a missing axis to None for fixed joints
and to [1,0,0] only in newer releases; the reference handles None at models/utils.py:228);
a joint without <limit> has limit None.
"""
import xml.etree.ElementTree as ET


def _floats(text, n=3):
    if text is None:
        return [0.0] * n
    return [float(v) for v in text.split()]


class Pose:
    def __init__(self, xyz=None, rpy=None):
        self.xyz = list(xyz) if xyz is not None else [0.0, 0.0, 0.0]
        self.rpy = list(rpy) if rpy is not None else [0.0, 0.0, 0.0]

    @property
    def position(self):
        return self.xyz

    @property
    def rotation(self):
        return self.rpy


class JointLimit:
    def __init__(self, effort=None, velocity=None, lower=None, upper=None):
        self.effort, self.velocity, self.lower, self.upper = effort, velocity, lower, upper


class JointDynamics:
    def __init__(self, damping=None, friction=None):
        self.damping, self.friction = damping, friction


class Inertia:
    def __init__(self, **kw):
        for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"):
            setattr(self, k, float(kw.get(k, 0.0)))


class Inertial:
    def __init__(self, mass=0.0, inertia=None, origin=None):
        self.mass, self.inertia, self.origin = mass, inertia, origin


class Box:
    def __init__(self, size=None):
        self.size = size


class Visual:
    def __init__(self, geometry=None, material=None, origin=None):
        self.geometry, self.material, self.origin = geometry, material, origin


class Collision:
    def __init__(self, geometry=None, origin=None):
        self.geometry, self.origin = geometry, origin


class Link:
    def __init__(self, name=None, visual=None, inertial=None, collision=None, origin=None):
        self.name, self.visual, self.inertial, self.collision, self.origin = name, visual, inertial, collision, origin


class Joint:
    def __init__(self, name=None, parent=None, child=None, joint_type=None, axis=None, origin=None,
                 limit=None, dynamics=None):
        self.name, self.parent, self.child, self.type = name, parent, child, joint_type
        self.axis, self.origin, self.limit, self.dynamics = axis, origin, limit, dynamics
        if self.origin is None:
            self.origin = Pose()

    @property
    def joint_type(self):
        return self.type


def _parse_origin(el):
    if el is None:
        return None
    return Pose(_floats(el.get("xyz")), _floats(el.get("rpy")))


class URDF:
    def __init__(self, name=None):
        self.name = name
        self.links, self.joints = [], []
        self.child_map, self.parent_map = {}, {}
        self.link_map, self.joint_map = {}, {}

    def add_link(self, link):
        self.links.append(link)
        self.link_map[link.name] = link

    def add_joint(self, joint):
        self.joints.append(joint)
        self.joint_map[joint.name] = joint
        self.parent_map[joint.child] = (joint.name, joint.parent)
        self.child_map.setdefault(joint.parent, []).append((joint.name, joint.child))

    @classmethod
    def from_xml_file(cls, path):
        root = ET.parse(str(path)).getroot()
        robot = cls(root.get("name"))
        for el in root:
            if el.tag == "link":
                inertial = None
                iel = el.find("inertial")
                if iel is not None:
                    mel, inel = iel.find("mass"), iel.find("inertia")
                    inertial = Inertial(
                        mass=float(mel.get("value")) if mel is not None else 0.0,
                        inertia=Inertia(**inel.attrib) if inel is not None else None,
                        origin=_parse_origin(iel.find("origin")))
                robot.add_link(Link(name=el.get("name"), inertial=inertial))
            elif el.tag == "joint":
                ax = el.find("axis")
                lim = el.find("limit")
                dyn = el.find("dynamics")
                limit = None
                if lim is not None:
                    limit = JointLimit(effort=float(lim.get("effort", 0.0)), velocity=float(lim.get("velocity", 0.0)),
                                       lower=float(lim.get("lower", 0.0)), upper=float(lim.get("upper", 0.0)))
                dynamics = None
                if dyn is not None:
                    dynamics = JointDynamics(damping=float(dyn.get("damping", 0.0)), friction=float(dyn.get("friction", 0.0)))
                origin = _parse_origin(el.find("origin"))
                robot.add_joint(Joint(
                    name=el.get("name"), parent=el.find("parent").get("link"), child=el.find("child").get("link"),
                    joint_type=el.get("type"), axis=_floats(ax.get("xyz")) if ax is not None else None,
                    origin=origin, limit=limit, dynamics=dynamics))
        return robot
