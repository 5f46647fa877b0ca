"""tezip_b200 -- B200-native TEZip predict-delta-encode hot path."""
__version__ = "0.1.0"
