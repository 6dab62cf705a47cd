"""ophelia_b200: B200-native dc_tts hot path (Text2Mel + SSRN) behind the operator surface of
CSTR-Edinburgh/ophelia's modules.py / networks.py / architectures.py.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
