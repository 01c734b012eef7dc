from .loss import Loss
from .fe_loss import FiniteElementLoss
from .mechanical import (MechanicalLoss, MechanicalLoss2DQuad, MechanicalLoss2DTri, MechanicalLoss3DHexa,
                         MechanicalLoss3DTetra)
from .thermal import ThermalLoss, ThermalLoss2DQuad, ThermalLoss2DTri, ThermalLoss3DHexa, ThermalLoss3DTetra
from .mechanical_neohooke import (NeoHookeMechanicalLoss, NeoHookeMechanicalLoss2DQuad,
                                  NeoHookeMechanicalLoss2DTri, NeoHookeMechanicalLoss3DHexa,
                                  NeoHookeMechanicalLoss3DTetra)
from .mechanical_elastoplasticity import (ElastoplasticityLoss, ElastoplasticityLoss2DQuad,
                                          ElastoplasticityLoss3DHexa, ElastoplasticityLoss3DTetra)
from .mechanical_saint_venant import (SaintVenantMechanicalLoss, SaintVenantMechanicalLoss2DQuad,
                                      SaintVenantMechanicalLoss2DTri, SaintVenantMechanicalLoss3DHexa,
                                      SaintVenantMechanicalLoss3DTetra)
from .transient_thermal import (TransientThermalLoss, TransientThermalLoss2DQuad, TransientThermalLoss2DTri,
                                TransientThermalLoss3DHexa, TransientThermalLoss3DTetra)
from .phase_field import AllenCahnLoss, AllenCahnLoss2DQuad, AllenCahnLoss2DTri, AllenCahnLoss3DHexa
from .kratos_small_displacement import KratosSmallDisplacement3DTetra
from .regression_loss import RegressionLoss
