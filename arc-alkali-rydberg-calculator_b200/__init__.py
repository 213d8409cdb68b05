"""arc_b200 -- B200-native (sm_100a) hot path of ARC (Alkali.ne Rydberg Calculator).

Drop-in for ARC's data-parallel path: radial matrix elements (Numerov), Hamiltonian
assembly and the per-sweep-point symmetric eigensolve behind the reference's own
Python API (`StarkMap`, `PairStateInteractions`, `getRadialMatrixElement`,
`NumerovWavefunction`).  All compute goes through libarcb200.so (hand-written CUDA);
there is no CPU fallback.
"""
from . import wigner  # noqa: F401
from .wigner import CG, Wigner3j, Wigner6j, WignerDmatrix  # noqa: F401
from .atoms import (AlkaliAtom, DivalentAtom, Hydrogen, Lithium6, Lithium7, Sodium,  # noqa: F401
                    Potassium, Potassium39, Potassium40, Potassium41, Rubidium,
                    Rubidium85, Rubidium87, Caesium, Cesium, Strontium88, Calcium40,
                    Ytterbium174)
from .numerov import NumerovWavefunction, NumerovBack  # noqa: F401
from .stark import StarkMap  # noqa: F401
from .shirley import StarkBasisGenerator, ShirleyMethod  # noqa: F401
from .pairstate import PairStateInteractions  # noqa: F401
from .resonances import StarkMapResonances  # noqa: F401

__version__ = "0.1.0"
