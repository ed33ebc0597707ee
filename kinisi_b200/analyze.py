"""The analyzer classes (mirror of kinisi/analyze.py:19-21)."""
from .conductivity_analyzer import ConductivityAnalyzer
from .diffusion_analyzer import DiffusionAnalyzer
from .jump_diffusion_analyzer import JumpDiffusionAnalyzer

__all__ = ['DiffusionAnalyzer', 'JumpDiffusionAnalyzer', 'ConductivityAnalyzer']
