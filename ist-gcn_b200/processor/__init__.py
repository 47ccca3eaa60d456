"""Drop-in for the reference ``processor`` package: ``processor.recognition.REC_Processor``."""
