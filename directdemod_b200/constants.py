"""Tunables and enumerations of the hot path.

The names and values are part of the drop-in contract (a decoder written against
directdemod/constants.py imports exactly these), so they are checked against the reference's
module in tests/test_host.py::test_constants_match_reference_values; how each one enters the GPU
path is noted next to it."""

# ---- enumerations (constants.py:28-36): small integers handed to source / filters.butter --------
SOURCE_IQWAV = 0
SOURCE_IQDAT = 1
FLT_LP = 0
FLT_HP = 1
FLT_BP = 2
FLT_BS = 3

# ---- names of the per-stream variables kept in the chunker (constants.py:39-40) -----------------
# "freqoffset": global sample index of the mixer (the fused chain's n0);
# "bwlim" + uniq: decimation phase carried from chunk to chunk (the fused chain's dec_off)
CHUNK_FREQOFFSET = "freqoffset"
CHUNK_BWLIM = "bwlim"

# ---- capture defaults (constants.py:4-8) ---------------------------------------------------------
IQ_SDRSAMPRATE = 2.048e6            # Hz, the RTL-SDR recordings of the reference's samples/
IQ_FREQOFFSET = 30000              # Hz the recordings are tuned off the carrier
PROC_CHUNKSIZE = 20000000         # samples per chunk: one fused launch each (160 MB of cf32)

# ---- NOAA APT (constants.py:11-23) ---------------------------------------------------------------
NOAA_FREQ = 137620000
NOAA_SATS = {137620000: "NOAA 15", 137100000: "NOAA 19", 137912500: "NOAA 18"}
NOAA_FMBW = 60000                  # FM bandwidth -> decimation 2 048 000 / 60 000 -> D = 34
NOAA_AUDSAMPRATE = 20800           # audio rate of getImage (strict FFT resample per chunk)
NOAA_CRUDESYNCSAMPRATE = 40960     # requested rate of the crude sync (non-strict: 60 235 Hz results)
NOAA_T = 1.0 / 4160                 # seconds per APT word
# sync words: four leading zeros, seven pulses (A: 2 on / 2 off at 1040 Hz, B: 3 on / 2 off at 832 Hz),
# zeros up to 40 words -- stretched to the sampling rate they become the run-length needles of ncc.cu
NOAA_SYNCA = [0, 0, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
NOAA_SYNCB = [0, 0, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 0]
NOAA_PEAKHEIGHTWIGGLE = 0.25        # threshold = top mean - this * (top mean - bottom mean)
NOAA_MINPEAKDIST = 0.45             # seconds between sync peaks: the window of ddm_pick_peaks
NOAA_DETECTMAXCHANGE = 5            # usefulness test: tolerated jitter of the 0.5 s sync spacing ...
NOAA_DETECTCONSSYNCSNUM = 10        # ... over this many consecutive syncs
NOAA_COLORCORRECT_FIFOLEN = 10000  # samples in the black / white level FIFOs of getImage
