# -*- coding: utf-8 -*-
"""popdynamics_b200: B200-native ensemble engine behind the popdynamics ``BaseModel`` API.

``popdynamics_b200.basepop`` is the drop-in for the reference module ``basepop``
(/root/reference/basepop.py); ``compiler`` / ``codegen`` lower a model to a flow table and CUDA C;
``capi`` binds the C-ABI library ``libpopdyn_b200.so`` (include/popdyn_b200.h); ``ensemble`` runs
batches of members on the GPU(s)."""

from . import basepop            # noqa: F401
from .basepop import BaseModel   # noqa: F401

__version__ = '0.1.0'
