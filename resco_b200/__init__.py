"""resco_b200 -- B200-native vectorised traffic-microsimulation backend behind RESCO's MultiSignal surface."""
__version__ = "0.1.0"
