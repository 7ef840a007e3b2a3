"""synference_b200: B200-native implementation of synference's mock-library hot path.

Public names follow ``src/synference/__init__.py:49-115`` of the reference for everything on the
path (prior draws, SFH basis, library builders, simulator, uncertainty models, hot-path utilities);
the inference / training half of the reference is out of scope and not re-exported.
"""

from .units import (Angstrom, Gyr, Jy, Msun, Myr, Quantity, Unit, mJy, nJy, uJy, um, unyt_array,  # noqa: F401
                    unyt_quantity, yr)
from .cosmology import FlatLambdaCDM, Planck18  # noqa: F401
from .parametric import (SFH, Blackbody, Calzetti2000, Greybody, EmergentEmission, EmissionModel, Filter, FilterCollection,  # noqa: F401
                         Grid, IncidentEmission, Instrument, IntrinsicEmission, PacmanEmission, BimodalPacmanEmission, PowerLaw,
                         SFHArray, TotalEmission, ZDist, ZDistArray)
from .igm import Inoue14  # noqa: F401
from .sampling import (continuity_sfh_array, draw_from_hypercube, generate_emission_models,  # noqa: F401
                       generate_metallicity_distribution, generate_sfh_basis, generate_sfh_grid, load_hypercube_from_npy)
from .utils import (asinh_err_to_f_jy, asinh_to_f_jy, asinh_to_snr, calculate_min_max_wav_grid,  # noqa: F401
                    check_log_scaling, check_scaling, f_jy_err_to_asinh, f_jy_to_asinh, generate_constant_R,
                    load_library_from_hdf5, combine_rank_files)
from .noise_models import (AsinhEmpiricalUncertaintyModel, DepthUncertaintyModel,  # noqa: F401
                           EmpiricalUncertaintyModel, GeneralEmpiricalUncertaintyModel,
                           SpectralUncertaintyModel, UncertaintyModel, create_uncertainty_models_from_EPOCHS_cat,
                           load_unc_model_from_hdf5, save_unc_model_to_hdf5)
from .engine import GalaxyParams, SynthEngine, depth_noise_features  # noqa: F401
from .library import CombinedBasis, GalaxyBasis, GalaxySimulator, create_galaxy  # noqa: F401
from .features import (ResampledFeatures, apply_empirical_noise_models,  # noqa: F401
                       create_feature_array_from_raw_photometry)
from .supplementary import (calculate_burstiness, calculate_mass_weighted_age, calculate_muv, calculate_MUV, calculate_sfh_quantile,  # noqa: F401
                            calculate_sfr, calculate_surviving_mass)
from .spectral import SpectrumResampler, create_feature_array_from_raw_spectra, transform_spectrum  # noqa: F401

__version__ = "0.1.0"
