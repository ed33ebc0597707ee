"""kinisi_b200: a B200-native (sm_100a) implementation of kinisi's displacement-statistics hot path
behind the kinisi.analyze API.  See DESIGN.md."""
__version__ = '0.1.0'
