"""Tunables and enums of the hot path; values are part of parity (directdemod/constants.py)."""

# IQ.wav settings (constants.py:4-5)
IQ_FREQOFFSET = 30000
IQ_SDRSAMPRATE = 2.048e6

# processing (constants.py:8)
PROC_CHUNKSIZE = 20000000

# NOAA (constants.py:11-23)
NOAA_FMBW = 60000
NOAA_AUDSAMPRATE = 20800
NOAA_FREQ = 137620000
NOAA_CRUDESYNCSAMPRATE = 40960
NOAA_T = 1.0 / 4160
NOAA_SYNCA = [0, 0, 0, 0] + [1, 1, 0, 0] * 7 + [0] * 8
NOAA_SYNCB = [0, 0, 0, 0] + [1, 1, 1, 0, 0] * 7 + [0]
NOAA_PEAKHEIGHTWIGGLE = 0.25
NOAA_MINPEAKDIST = 0.45
NOAA_COLORCORRECT_FIFOLEN = 10000
NOAA_DETECTMAXCHANGE = 5
NOAA_DETECTCONSSYNCSNUM = 10
NOAA_SATS = {137620000: "NOAA 15", 137100000: "NOAA 19", 137912500: "NOAA 18"}

# source types (constants.py:28-29)
SOURCE_IQWAV = 0
SOURCE_IQDAT = 1

# filter types (constants.py:33-36)
FLT_LP = 0
FLT_HP = 1
FLT_BP = 2
FLT_BS = 3

# chunker variable names (constants.py:39-40)
CHUNK_FREQOFFSET = "freqoffset"
CHUNK_BWLIM = "bwlim"
