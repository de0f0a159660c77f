"""Minimal stand-in for the part of MDAnalysis that JeLLyFysh's PdbInputHandler uses
(jellyfysh/input_output_handler/input_handler/pdb_input_handler.py:107-190 via mdanalysis_import.py:33), for
installations without MDAnalysis: `Universe(filename)` of a .pdb file with

    universe.dimensions            [a, b, c, alpha, beta, gamma] of the CRYST1 record
    universe.atoms                 atoms in file order: .id (serial), .resid, .name, .position (numpy float32[3])
    universe.residues              residues in order of first appearance: .resid, .atoms

Coordinates are kept as float32, as MDAnalysis keeps them: the reference converts them with float(), so the start
configuration of a run is the float32-rounded content of the file (caveat 1).
`jellyfysh_b200.install()` puts this package on the import path only when the real MDAnalysis is not importable.
Writing trajectories (PdbOutputHandler, DcdOutputHandler) is not covered: `Writer` says so."""
import numpy as np

__version__ = "0-jellyfysh-b200-shim"


class _Atom:
    def __init__(self, serial, name, resname, resid, position):
        self.id, self.name, self.resname, self.resid = serial, name, resname, resid
        self.position = position


class _Residue:
    def __init__(self, resid):
        self.resid = resid
        self.atoms = []


class Universe:
    def __init__(self, filename, *args, **kwargs):
        if args or kwargs:
            raise NotImplementedError("the MDAnalysis stand-in of jellyfysh_b200 only reads one .pdb file")
        self.filename = filename
        self.atoms, self.residues = [], []
        self.dimensions = np.zeros(6, dtype=np.float32)
        by_resid = {}
        with open(filename) as handle:
            for line in handle:
                record = line[:6]
                if record == "CRYST1":
                    self.dimensions = np.array([line[6:15], line[15:24], line[24:33], line[33:40], line[40:47], line[47:54]],
                                               dtype=np.float32)
                elif record in ("ATOM  ", "HETATM"):
                    resid = int(line[22:26])
                    position = np.array([line[30:38], line[38:46], line[46:54]], dtype=np.float32)
                    atom = _Atom(int(line[6:11]), line[12:16].strip(), line[17:21].strip(), resid, position)
                    self.atoms.append(atom)
                    if resid not in by_resid:
                        by_resid[resid] = _Residue(resid)
                        self.residues.append(by_resid[resid])
                    by_resid[resid].atoms.append(atom)

    @classmethod
    def empty(cls, *args, **kwargs):
        raise NotImplementedError("the MDAnalysis stand-in of jellyfysh_b200 does not build universes for writing")


class Writer:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the MDAnalysis stand-in of jellyfysh_b200 reads .pdb start configurations only; "
                                  "install MDAnalysis for PdbOutputHandler / DcdOutputHandler")
