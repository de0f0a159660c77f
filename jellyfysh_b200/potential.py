"""Host-side Potential classes over the device arithmetic (SURVEY.md 8b "Potential contract").

The classes carry the reference's names and constructor arguments and keep its calling convention --
`derivative(velocity, separation, *charges)` and `displacement(velocity, separation, *charges, potential_change)`
(jellyfysh/potential/potential.py:154-181, 218-301), with the properties `number_separation_arguments`,
`number_charge_arguments` and `potential_change_required` the event handlers introspect -- but every number comes from
libecmc_b200's `ecmc_potential_derivative` / `ecmc_potential_displacement` (csrc/ecmc_math.cuh on the device), so the
reference's own unit tests of a potential can be pointed at the code the event kernels run (tests/test_gpu_potentials.py
replays their known answers through these classes). `derivatives(...)` / `displacements(...)` are the batched forms: one
launch for n separations.

The periodic box is an argument here (`system_length`, `dimension`; default: the values of the reference's
`jellyfysh.setting` when that package is importable and initialised), because the merged-image Coulomb sum and the
bounding potential depend on it.
"""
import numpy as np

from jellyfysh_b200 import abi, engine


def _box(system_length, dimension):
    if system_length is not None and dimension is not None:
        return float(system_length), int(dimension)
    try:  # the reference's global setting (jellyfysh/setting/__init__.py), if it is there
        from jellyfysh import setting
        from jellyfysh.setting import hypercubic_setting
        return (float(hypercubic_setting.system_length) if system_length is None else float(system_length),
                int(setting.dimension) if dimension is None else int(dimension))
    except Exception as error:  # noqa: BLE001
        raise ValueError("system_length and dimension are needed (no initialised jellyfysh.setting)") from error


class DevicePotential:
    """Base: an EcmcPotential record plus the box; subclasses set the reference's introspection properties."""
    number_separation_arguments = 1
    number_charge_arguments = 0
    potential_change_required = True

    def __init__(self, record, system_length=None, dimension=None, device=0):
        self._record = record
        self._system_length, self._dimension = system_length, dimension
        self._device = int(device)

    @property
    def record(self):
        """The EcmcPotential (include/ecmc.h) an EcmcProgram takes."""
        return self._record

    def _charges(self, n, charges):
        if self.number_charge_arguments == 0:
            return None
        if len(charges) != self.number_charge_arguments:
            raise TypeError("{0} takes {1} charges".format(type(self).__name__, self.number_charge_arguments))
        return np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), (n,)) for c in charges], axis=1)

    def derivatives(self, velocity, separations, *charges):
        """Batched Potential.derivative: separations[n][dimension], charges scalars or [n] each."""
        length, dimension = _box(self._system_length, self._dimension)
        separations = np.asarray(separations, dtype=np.float64).reshape(-1, dimension)
        return engine.potential_derivative(self._record, dimension, length, velocity, separations,
                                           self._charges(len(separations), charges), device=self._device)

    def displacements(self, velocity, separations, *args):
        """Batched InvertiblePotential.displacement: args = charges..., then potential changes if required."""
        length, dimension = _box(self._system_length, self._dimension)
        separations = np.asarray(separations, dtype=np.float64).reshape(-1, dimension)
        charges = args[:self.number_charge_arguments]
        rest = args[self.number_charge_arguments:]
        changes = None
        if self.potential_change_required:
            if len(rest) != 1:
                raise TypeError("{0}.displacement needs the potential change".format(type(self).__name__))
            changes = np.broadcast_to(np.asarray(rest[0], dtype=np.float64), (len(separations),))
        return engine.potential_displacement(self._record, dimension, length, velocity, separations,
                                             self._charges(len(separations), charges), changes, device=self._device)

    def derivative(self, velocity, separation, *charges):
        """Potential.derivative (potential.py:154-181): the time derivative for one separation."""
        return float(self.derivatives(velocity, [separation], *charges)[0])

    def displacement(self, velocity, separation, *args):
        """InvertiblePotential.displacement (potential.py:218-301): the time until the potential change is reached."""
        return float(self.displacements(velocity, [separation], *args)[0])


class InversePowerPotential(DevicePotential):
    """c1 c2 k / r^p (inverse_power_potential.py:43-179)."""
    number_charge_arguments = 2

    def __init__(self, power, prefactor, **box):
        super().__init__(abi.EcmcPotential.make(abi.POT_INVERSE_POWER, float(power), float(prefactor)), **box)


class LennardJonesPotential(DevicePotential):
    """k ((s / r)^12 - (s / r)^6) (lennard_jones_potential.py:43-135 on potential/abstracts.py:336-530)."""

    def __init__(self, prefactor=1.0, characteristic_length=1.0, **box):
        super().__init__(abi.EcmcPotential.make(abi.POT_LENNARD_JONES, float(prefactor), float(characteristic_length)), **box)


class DisplacedEvenPowerPotential(DevicePotential):
    """k (r - r0)^p, p even (displaced_even_power_potential.py:44-142)."""

    def __init__(self, equilibrium_separation, power, prefactor=1.0, **box):
        super().__init__(abi.EcmcPotential.make(abi.POT_DISPLACED_EVEN_POWER, float(prefactor),
                                                float(equilibrium_separation), float(power)), **box)


class HardSpherePotential(DevicePotential):
    """Hard spheres of one radius (hard_sphere_potential.py:44-99); any velocity."""
    potential_change_required = False

    def __init__(self, radius, **box):
        super().__init__(abi.EcmcPotential.make(abi.POT_HARD_SPHERE, float(radius)), **box)


class HardDipolePotential(DevicePotential):
    """Hard tether between a minimum and a maximum separation (hard_dipole_potential.py:42-114); any velocity."""
    potential_change_required = False

    def __init__(self, minimum_separation, maximum_separation, **box):
        super().__init__(abi.EcmcPotential.make(abi.POT_HARD_DIPOLE, float(minimum_separation), float(maximum_separation)),
                         **box)


class MergedImageCoulombPotential(DevicePotential):
    """Ewald-summed Coulomb interaction of all periodic images (merged_image_coulomb_potential.c:77-274); derivative only."""
    number_charge_arguments = 2
    potential_change_required = False

    def __init__(self, alpha=3.45, fourier_cutoff=6, position_cutoff=2, prefactor=1.0, **box):
        super().__init__(abi.EcmcPotential.make(abi.POT_MERGED_IMAGE_COULOMB, float(prefactor), float(alpha),
                                                float(fourier_cutoff), float(position_cutoff)), **box)

    def displacement(self, *args):
        raise NotImplementedError("the merged-image Coulomb potential is not invertible (the reference bounds it)")

    displacements = displacement


class InversePowerCoulombBoundingPotential(DevicePotential):
    """Bounding potential of the merged-image Coulomb potential (inverse_power_coulomb_bounding_potential.c:53-139)."""
    number_charge_arguments = 2

    def __init__(self, prefactor=1.5837, **box):
        super().__init__(abi.EcmcPotential.make(abi.POT_INVERSE_POWER_COULOMB_BOUNDING, float(prefactor)), **box)


BY_NAME = {cls.__name__: cls for cls in (InversePowerPotential, LennardJonesPotential, DisplacedEvenPowerPotential,
                                         HardSpherePotential, HardDipolePotential, MergedImageCoulombPotential,
                                         InversePowerCoulombBoundingPotential)}
