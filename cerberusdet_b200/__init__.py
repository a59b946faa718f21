"""cerberusdet_b200 -- B200-native (sm_100a) post-head path for CerberusDet."""
__version__ = "0.1.0"
