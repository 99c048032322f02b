"""A small molecular data model with the accessors KREG uses on ``mlatom.data`` objects
(reference: mlatom/data.py -- ``molecular_database.xyz_coordinates`` :2320-2328, ``get_properties`` :2075,
``get_xyz_vectorial_properties`` :2098, ``add_scalar_properties`` :1857, ``add_xyz_vectorial_properties``
:1920).  The model classes are duck-typed: real ``mlatom.data.molecule`` / ``molecular_database``
objects work unchanged; these classes exist so the package is usable (and testable on the GPU box)
without the reference installed, and to give array-backed construction for large batches."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np


class atom:
    def __init__(self, atomic_number: int = 0, xyz_coordinates=None):
        self.atomic_number = int(atomic_number)
        self.xyz_coordinates = np.zeros(3) if xyz_coordinates is None else np.asarray(xyz_coordinates, dtype=float)


class molecule:
    def __init__(self, atoms: Optional[List[atom]] = None):
        self.atoms = atoms if atoms is not None else []

    @classmethod
    def from_arrays(cls, atomic_numbers: Sequence[int], xyz) -> "molecule":
        xyz = np.asarray(xyz, dtype=float)
        return cls([atom(z, xyz[i]) for i, z in enumerate(atomic_numbers)])

    @property
    def xyz_coordinates(self) -> np.ndarray:
        return np.array([a.xyz_coordinates for a in self.atoms])

    @xyz_coordinates.setter
    def xyz_coordinates(self, value):
        for a, v in zip(self.atoms, np.asarray(value, dtype=float)):
            a.xyz_coordinates = v

    @property
    def atomic_numbers(self) -> np.ndarray:
        return np.array([a.atomic_number for a in self.atoms])

    def get_property(self, name):
        return self.__dict__[name]

    def add_scalar_property(self, scalar, property_name="y"):
        self.__dict__[property_name] = scalar

    def add_xyz_vectorial_property(self, vector, xyz_vectorial_property="xyz_vector"):
        for a, v in zip(self.atoms, vector):
            a.__dict__[xyz_vectorial_property] = v

    add_xyz_derivative_property = add_xyz_vectorial_property

    def get_xyz_vectorial_properties(self, name) -> np.ndarray:
        return np.array([a.__dict__[name] for a in self.atoms])

    def get_energy_gradients(self):
        return self.get_xyz_vectorial_properties("energy_gradients")

    def get_internuclear_distance_matrix(self) -> np.ndarray:
        x = self.xyz_coordinates
        return np.linalg.norm(x[:, None, :] - x[None, :, :], axis=-1)

    def copy(self) -> "molecule":
        import copy

        return copy.deepcopy(self)


class molecular_database:
    def __init__(self, molecules: Optional[Iterable[molecule]] = None):
        self.molecules = list(molecules) if molecules is not None else []

    @classmethod
    def from_arrays(cls, atomic_numbers, xyz, **properties) -> "molecular_database":
        """xyz (N, Nat, 3); scalar properties as (N,), xyz-vectorial ones as (N, Nat, 3)."""
        xyz = np.asarray(xyz, dtype=float)
        db = cls([molecule.from_arrays(atomic_numbers, g) for g in xyz])
        for name, values in properties.items():
            values = np.asarray(values)
            if values.ndim == 1:
                db.add_scalar_properties(values, name)
            else:
                db.add_xyz_vectorial_properties(values, name)
        return db

    def __len__(self):
        return len(self.molecules)

    def __getitem__(self, item):
        if isinstance(item, slice):
            return molecular_database(self.molecules[item])
        return self.molecules[item]

    def __iter__(self):
        return iter(self.molecules)

    @property
    def xyz_coordinates(self) -> np.ndarray:
        return np.array([m.xyz_coordinates for m in self.molecules])

    def get_properties(self, property_name="y") -> np.ndarray:
        return np.array([m.get_property(property_name) for m in self.molecules])

    def get_xyz_vectorial_properties(self, property_name) -> np.ndarray:
        return np.array([m.get_xyz_vectorial_properties(property_name) for m in self.molecules])

    get_xyz_derivative_properties = get_xyz_vectorial_properties

    def add_scalar_properties(self, scalars, property_name="y") -> None:
        for i in range(len(scalars)):
            self.molecules[i].add_scalar_property(scalars[i], property_name=property_name)

    def add_xyz_vectorial_properties(self, vectors, xyz_vectorial_property="xyz_vector") -> None:
        for i in range(len(vectors)):
            self.molecules[i].add_xyz_vectorial_property(vectors[i], xyz_vectorial_property=xyz_vectorial_property)

    add_xyz_derivative_properties = add_xyz_vectorial_properties

    # ---- text files (mlatom/data.py:1568-1578, 1869-1882, 1944-1990, 2120-2130), through the array parsers ----
    def read_from_xyz_file(self, filename: str, append: bool = False) -> "molecular_database":
        from . import textio

        z, xyz = textio.read_xyz(filename)
        mols = [molecule.from_arrays(z[i], xyz[i]) for i in range(len(z))]
        self.molecules = (self.molecules + mols) if append else mols
        return self

    @classmethod
    def from_xyz_file(cls, filename: str) -> "molecular_database":
        return cls().read_from_xyz_file(filename)

    def add_scalar_properties_from_file(self, filename: str, property_name: str = "y") -> None:
        from . import textio

        self.add_scalar_properties(textio.read_y(filename), property_name=property_name)

    def add_xyz_vectorial_properties_from_file(self, filename: str, xyz_vectorial_property: str = "xyz_vector") -> None:
        from . import textio

        self.add_xyz_vectorial_properties(textio.read_xyz_vectors(filename), xyz_vectorial_property=xyz_vectorial_property)

    add_xyz_derivative_properties_from_file = add_xyz_vectorial_properties_from_file

    def write_file_with_xyz_coordinates(self, filename: str) -> None:
        from . import textio

        z = np.array([m.atomic_numbers for m in self.molecules])
        textio.write_xyz(filename, z, self.xyz_coordinates)

    def write_file_with_properties(self, filename: str, property_to_write: str = "y") -> None:
        from . import textio

        textio.write_y(filename, self.get_properties(property_to_write))

    def write_file_with_xyz_vectorial_properties(self, filename: str, xyz_vectorial_property_to_write: str = "xyz_vector") -> None:
        from . import textio

        textio.write_xyz_vectors(filename, self.get_xyz_vectorial_properties(xyz_vectorial_property_to_write))

    def write_file_with_xyz_derivative_properties(self, filename: str, xyz_derivative_property_to_write: str = "xyz_derivatives") -> None:
        self.write_file_with_xyz_vectorial_properties(filename, xyz_derivative_property_to_write)

    def write_file_energy_gradients(self, filename: str) -> None:
        self.write_file_with_xyz_vectorial_properties(filename, "energy_gradients")

