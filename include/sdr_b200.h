/* include/sdr_b200.h -- C-ABI of the B200 SDR DSP engine (libsdr_b200.so).
 *
 * Drop-in boundary for CubicSDR's spectrum + demodulator hot path.  The reference has no FFI on this path; its
 * host code speaks liquid-dsp's C API (opaque handle, create / execute / destroy, int status, caller-owned arrays,
 * "num written" out-parameters).  This ABI keeps that shape, so the three in-process seams of the reference can
 * bind to it with a few lines each (INTEGRATION.md shows the bindings).  File:line citations are relative to
 * /root/reference/.
 *
 * Threading: calls on one engine are serialised internally; different engines are independent.
 * Errors: every function returns 0 on success, a negative sdr_status otherwise; sdr_last_error() gives text.
 * All sample data is float32, complex = interleaved {re, im} (liquid_float_complex, liquid.h:153).
 */
#ifndef SDR_B200_H
#define SDR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdr_engine sdr_engine;       /* one per GPU: owns streams, the wideband block buffer, all channels */
typedef struct sdr_chan sdr_chan;           /* one demodulator channel (DemodulatorPreThread + Modem + metering)  */
typedef struct sdr_spectrum sdr_spectrum;   /* one SpectrumVisualProcessor-equivalent                              */
typedef struct { float re, im; } sdr_cf32;

enum sdr_status {
    SDR_OK = 0, SDR_ERR_ARG = -1, SDR_ERR_CUDA = -2, SDR_ERR_NOMEM = -3, SDR_ERR_STATE = -4, SDR_ERR_UNSUPPORTED = -5
};

const char *sdr_last_error(void);
const char *sdr_version(void);
/* 1 when the library was built for the GPU (always, for libsdr_b200.so); the test-only emulator reports 0 */
int sdr_is_device_build(void);

/* ---- engine ------------------------------------------------------------------------------------------------------ */
/* device: CUDA ordinal.  max_block: largest block (complex samples) sdr_engine_submit will accept.
 * stream: a cudaStream_t to run on (as void*), or NULL to create one.                                              */
int sdr_engine_create(int device, unsigned max_block, void *stream, sdr_engine **out);
int sdr_engine_destroy(sdr_engine *eng);

/* Replaces SDRPostThread::runSingleCH (src/sdr/SDRPostThread.cpp:248-299): takes one block of raw IQ, applies the
 * DC blocker (:284), and runs every attached spectrum and channel on it.  Asynchronous: returns once the work is
 * enqueued; results are fetched per channel / spectrum after sdr_engine_wait() / sdr_engine_wait_block().  center_freq /
 * sample_rate are the SDRThreadIQData fields (src/sdr/SoapySDRThread.h:20-39).  iq is host memory.
 * Like the reference's post thread, which only try_push()es (:238-298), the call never waits for the GPU in steady state:
 * a pageable block is copied once into a ring of three pinned buffers; a block that already lies in pinned memory (the
 * engine's own ring buffer from sdr_engine_host_buffer, sdr_host_alloc memory, anything cudaHostRegister'ed) is sent by
 * DMA from where it lies and must then stay untouched until two further blocks have been submitted or a wait call has
 * returned for it.                                                                                                    */
int sdr_engine_submit(sdr_engine *eng, const sdr_cf32 *iq, unsigned n, long long center_freq, long long sample_rate);
/* The pinned buffer the NEXT sdr_engine_submit takes without a staging copy (capacity max_block samples): what a producer
 * fills in place of SDRThreadIQData::data (SDRThread::readStream appends the stream to that vector, src/sdr/SoapySDRThread.cpp:195-380).   */
int sdr_engine_host_buffer(sdr_engine *eng, sdr_cf32 **buf, unsigned *cap);
/* Pinned host memory for blocks and result arrays that live on the host side of the boundary. */
int sdr_host_alloc(size_t bytes, void **out);
int sdr_host_free(void *p);
/* Same, for a block that already lives in this GPU's memory (e.g. after an NCCL broadcast).  No host copy is made;
 * with fetch_to_host = 0 results stay in device memory (throughput measurement), otherwise they are copied back.  */
int sdr_engine_submit_device(sdr_engine *eng, const void *d_iq, unsigned n, long long center_freq, long long sample_rate,
                             int fetch_to_host);
/* Pipelined form: the block is ready when `ready_event` (a cudaEvent_t, as void*) fires, or already is if it is NULL --
 * independent of what is queued on the engine's stream.  This lets the DC blocker of block b+1 run underneath the
 * channel kernels of block b.  The caller must not overwrite d_iq before a later sdr_engine_wait().              */
int sdr_engine_submit_device_async(sdr_engine *eng, const void *d_iq, unsigned n, long long center_freq, long long sample_rate,
                                   int fetch_to_host, void *ready_event);
int sdr_engine_wait(sdr_engine *eng);
/* Pipelined consumer (the reference's threads run blocks apart, ThreadBlockingQueue between SDRPostThread.cpp:294-298
 * and DemodulatorPreThread.cpp:53): wait only until the host results of the block submitted `lag` submits ago (0 = the
 * latest, 1, or 2) have landed, and point the fetch calls at that block.  submit(b+2); wait_block(2); fetch = block b,
 * while b+1 and b+2 are still on the GPU.  Host results sit in a ring of three buffers, so at most two newer blocks may
 * be in flight when a block is fetched. */
int sdr_engine_wait_block(sdr_engine *eng, int lag);
/* DC-blocked block of the last submit (what runSingleCH pushes to the visual queues / demod pipes), host copy.    */
int sdr_engine_fetch_block(sdr_engine *eng, sdr_cf32 *out, unsigned cap, unsigned *n);
/* kernel launches issued by this engine so far (bench.py's gpu_launches) */
unsigned long long sdr_engine_launch_count(sdr_engine *eng);
/* Per-stage device time, measured with CUDA events on the engine's stream around each stage of every submit while
 * enabled (bench.py's roofline figure).  stage_ms4 = accumulated {DC blocker, spectrum, channel front kernel(s),
 * channel tail kernel} milliseconds over `blocks` submits; collected at sdr_engine_wait (at most 64 submits per wait). */
int sdr_engine_set_timing(sdr_engine *eng, int enable);
int sdr_engine_get_timing(sdr_engine *eng, double *stage_ms4, unsigned long long *blocks);

/* ---- channel: DemodulatorPreThread + Modem + DemodulatorThread metering --------------------------------------------- */
/* modem: "FM" "NBFM" "AM" "CW" "USB" "LSB" "DSB" "I/Q" (Modem::addModemFactory names, src/CubicSDR.cpp:305-313).
 * bandwidth is passed through Modem::checkSampleRate (ModemAnalog.cpp:14-19, ModemUSB.cpp:29-37).
 * Builds what DemodulatorWorkerThread builds (src/demod/DemodulatorWorkerThread.cpp:63-101): msresamp_crcf(bw/fs, 60),
 * the modem, and its kit msresamp_rrrf(audio_rate/bw, 60) (ModemAnalog.cpp:21-33).
 * Like DemodulatorWorkerThread, create / destroy may run on any thread while another one submits blocks: they never make
 * the submitter wait (device memory comes from the stream-ordered allocator; a new channel joins with the next block, a
 * destroyed one is recycled once the blocks in flight have drained).  A destroyed handle must not be used again, and
 * results of blocks submitted before the destroy can no longer be fetched for it.                                      */
int sdr_chan_create(sdr_engine *eng, const char *modem, long long sample_rate, long long bandwidth, int audio_rate,
                    sdr_chan **out);
int sdr_chan_destroy(sdr_chan *ch);
/* Demodulator frequency (absolute, Hz).  Shift = frequency - block centre is applied per block exactly as
 * DemodulatorPreThread.cpp:154-163 does, including abs(), the up/down choice and the 0.75 fs skip rule.            */
int sdr_chan_set_frequency(sdr_chan *ch, long long frequency);
int sdr_chan_get_bandwidth(sdr_chan *ch, int *bandwidth);
/* AudioThreadInput::channels: 2 for "I/Q" (audio = interleaved {imag, real} of the decimated IQ, ModemIQ.cpp:46-55), else 1 */
int sdr_chan_audio_channels(sdr_chan *ch, int *channels);
/* most audio samples / decimated IQ samples one block can yield for this channel (size the fetch buffers with these) */
int sdr_chan_result_caps(sdr_chan *ch, unsigned *audio_cap, unsigned *iq_cap);
/* also return the decimated IQ of each block (DemodulatorThreadPostIQData::data, src/demod/DemodDefs.h:43-59) */
int sdr_chan_want_iq(sdr_chan *ch, int enable);
/* Results of the last submitted block (after sdr_engine_wait): audio = AudioThreadInput::data
 * (src/audio/AudioThread.h:16-51), level_db = DemodulatorThread.cpp:142-163, peak = :223-233.  skipped = 1 when the
 * 0.75 fs rule dropped the block for this channel.                                                                */
int sdr_chan_fetch_audio(sdr_chan *ch, float *audio, unsigned cap, unsigned *n_audio, double *level_db, float *peak,
                         int *skipped);
int sdr_chan_fetch_iq(sdr_chan *ch, sdr_cf32 *iq, unsigned cap, unsigned *n_iq);
/* DemodulatorThread's squelch (src/demod/DemodulatorThread.cpp:165-198, 384-397): the floor / ceiling trackers, the
 * smoothed signal level and the squelch decision are advanced on the device with every block the channel processes.
 * sdr_chan_set_squelch = setSquelchEnabled + setSquelchLevel (dB), effective from the next block.  sdr_chan_fetch_meter
 * returns getSignalLevel / getSignalFloor / getSignalCeil as they stand after the fetched block and that block's
 * AudioThreadInput::is_squelch_active (the host drops the audio of a squelched block, :318).                         */
int sdr_chan_set_squelch(sdr_chan *ch, int enabled, float level_db);
int sdr_chan_fetch_meter(sdr_chan *ch, float *signal_level, float *signal_floor, float *signal_ceil, int *squelched);
/* One call for many channels (what a host with hundreds of demodulators uses instead of n_chan fetches): audio of
 * channel i lands at audio + i * stride (at most stride samples); the other arrays have n_chan entries and may be NULL. */
int sdr_engine_fetch_audio_batch(sdr_engine *eng, sdr_chan *const *chans, unsigned n_chan, float *audio, unsigned stride,
                                 unsigned *n_audio, double *level_db, float *peak);

/* Zero-copy form: audio[i] points into the pinned host buffer the block's results were DMA'd to (n_audio[i] floats).  The
 * pointers stay valid until two further blocks have been submitted.  For hosts that pass the samples on at once.         */
int sdr_engine_view_audio_batch(sdr_engine *eng, sdr_chan *const *chans, unsigned n_chan, const float **audio,
                                unsigned *n_audio, double *level_db, float *peak);

/* ---- Modem plugin seam: one call = Modem::demodulate(kit, input, audioOut) (src/modules/modem/Modem.h:148-153) -------- */
/* Modem::checkSampleRate for `modem` (ModemAnalog.cpp:14-19, ModemUSB.cpp:29-37); negative on unknown modem.       */
int sdr_modem_check_sample_rate(const char *modem, long long sample_rate);
/* both arguments of Modem::checkSampleRate(sampleRate, audioSampleRate): "I/Q" returns the audio rate (ModemIQ.cpp:31-33) */
int sdr_modem_check_sample_rate2(const char *modem, long long sample_rate, int audio_rate);
/* Modem factory + buildKit(sample_rate, audio_rate) (Modem.cpp:40-55, ModemAnalog.cpp:21-33).  The handle is an sdr_chan
 * without the pre-demod stage: sdr_engine_submit ignores it; destroy it with sdr_chan_destroy (= disposeKit + delete). */
int sdr_modem_create(sdr_engine *eng, const char *modem, long long sample_rate, int audio_rate, unsigned max_samples,
                     sdr_chan **out);
/* iq: n samples already at the modem rate (DemodulatorThreadPostIQData::data).  Synchronous.  Also returns what
 * DemodulatorThread computes right after the call: level_db (:142-163) and peak (:223-233).  n == 0 is a no-op.   */
int sdr_modem_demodulate(sdr_chan *modem, const sdr_cf32 *iq, unsigned n, float *audio, unsigned cap, unsigned *n_audio,
                         double *level_db, float *peak);

/* ---- spectrum: FFTDataDistributor framing + SpectrumVisualProcessor + WaterfallPanel::step --------------------------- */
/* fft_size: display points (SpectrumVisualProcessor::setup, :140-178); the transform is 2 * fft_size.
 * lines_per_second: FFTDataDistributor::setLinesPerSecond; 0 = transform every non-overlapping frame.             */
int sdr_spectrum_create(sdr_engine *eng, unsigned fft_size, unsigned lines_per_second, sdr_spectrum **out);
int sdr_spectrum_destroy(sdr_spectrum *sp);
int sdr_spectrum_set_average_rate(sdr_spectrum *sp, float rate);      /* setFFTAverageRate, default 0.65 (:36) */
/* Zoomed view: SpectrumVisualProcessor::setView(bView, centerFreq, bandwidth) (src/process/SpectrumVisualProcessor.cpp:57-75;
 * setCenterFrequency / setBandwidth :87-110 are the same call with one argument changed).  In view mode every frame is
 * shifted by the view's oscillator and decimated by msresamp_crcf(resampleBw / sampleRate) -- resampleBw = the smallest
 * power-of-two fraction of the rate that still holds the bandwidth -- before the transform (:283-386); on a zoom step or a
 * pan the running averages are re-binned / moved (:317-329, :454-492) and the display points span bandwidth / resampleBw of
 * the bins (:532-556).  Effective from the next block.  A frame of the view is fft_size * 2 * sampleRate / resampleBw
 * samples (sdr_spectrum_desired_input = getDesiredInputSize) and must fit the engine's stream history (one block).     */
int sdr_spectrum_set_view(sdr_spectrum *sp, int is_view, long long center_freq, long bandwidth);
int sdr_spectrum_desired_input(sdr_spectrum *sp, long long sample_rate, unsigned *n);
/* frames produced by the last submitted block */
int sdr_spectrum_frames(sdr_spectrum *sp, unsigned *n_frames);
/* frame `idx` of the last block: points_y = SpectrumVisualData::spectrum_points y values (fft_size floats),
 * line = the u8 waterfall line, ceil_floor = {fft_ceil_maa, fft_floor_maa}.  Any pointer may be NULL.            */
int sdr_spectrum_fetch(sdr_spectrum *sp, unsigned idx, float *points_y, uint8_t *line, double *ceil_floor);
/* zero-copy: all frames of the fetched block, [n_frames][fft_size] each, in the pinned result buffer (same lifetime as above) */
int sdr_spectrum_view(sdr_spectrum *sp, const float **points_y, const uint8_t **lines, unsigned *n_frames);
/* every frame of the last block in one call: points_y [cap_frames][fft_size], lines likewise, ceil_floor [cap_frames][2] */
int sdr_spectrum_fetch_all(sdr_spectrum *sp, unsigned cap_frames, float *points_y, uint8_t *lines, double *ceil_floor,
                           unsigned *n_frames);
/* diagnostics for parity tests: linear fftshifted magnitudes and float(maa) of frame idx (2 * fft_size floats each) */
int sdr_spectrum_fetch_debug(sdr_spectrum *sp, unsigned idx, float *mags, float *maa);

/* ---- polyphase filterbank channelizer -----------------------------------------------------------
 * Replaces firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, numChannels, 4, 60) (src/sdr/SDRPostThread.cpp:406) and the
 * per-numChannels-samples firpfbch_crcf_analyzer_execute loop of SDRPostThread::runPFBCH (:447-451).  Output layout is
 * liquid's: numChannels interleaved channel samples per input frame, which runDemodChannels de-interleaves (:366-382).
 * n must be a multiple of num_channels (SDRThread::getOptimalElementCount, src/sdr/SoapySDRThread.cpp:668-674). */
/* Channelizer mode of the engine: SDRPostThread::setChannelizerType / the dispatch of SDRPostThread::run (:142-148,
 * :178-187).  mode 0 (default) = runSingleCH for every block; mode 2 = SDRPostPFBCH2 (as mode 1 with firpfbch2: the
 * demodulators run at twice the channel rate); mode 1 = SDRPostPFBCH: a block whose rate gives
 * numChannels > 1 goes through the analysis filterbank (no wideband DC blocker, :432-451), demodulators are routed to
 * the nearest channel centre (updateChannels / getChannelAt / runDemodChannels, :115-139, :303-398; channel 0 gets the DC
 * blocker, the extra channel wraps onto numChannels/2) and must have been created with sample_rate = channel_rate of
 * sdr_engine_channel_plan (what DemodulatorPreThread sees as inp->sampleRate).  Spectra see the raw block, as there. */
int sdr_engine_set_channelizer(sdr_engine *eng, int mode);
int sdr_engine_channel_plan(long long sample_rate, int *num_channels, long long *channel_rate);
/* The wideband DC blocker (iirfilt_crcf_execute_block, src/sdr/SDRPostThread.cpp:284) is replayed bit for bit in one of two
 * ways: 1 = candidate windows (no warm-up: the lowest latency, some chip-wide work), 0 = warmed-up segments on a few reserved
 * SMs (hides under a long channel front kernel), -1 (default) = chosen per block from the number of attached channels.
 * Both give identical samples; this only moves time around.  Environment: SDR_B200_DC_WINDOW=0|1. */
int sdr_engine_set_dc_window(sdr_engine *eng, int mode);
/* Most samples one demodulator receives per block (channelizer mode: block / num_channels, twice that in mode 2); sizes the
 * channels' buffers and result slots.  Default max_block; call before creating channels. */
int sdr_engine_set_channel_block_max(sdr_engine *eng, unsigned n);
typedef struct sdr_pfb sdr_pfb;
int sdr_pfb_optimal_channel_count(long long sample_rate);     /* SDRThread::getOptimalChannelCount, SoapySDRThread.cpp:676-693 */
int sdr_pfb_create(sdr_engine *eng, unsigned num_channels, unsigned m, float as_db, unsigned max_frames, sdr_pfb **out);
/* firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, numChannels, 4, 60) + firpfbch2_crcf_execute per numChannels/2 samples
 * (SDRPostThread::initPFBCH2 / runPFBCH2, SDRPostThread.cpp:457-511): 2 n output samples for n inputs; max_frames counts
 * output frames.  Same execute / reset / destroy calls. */
int sdr_pfb2_create(sdr_engine *eng, unsigned num_channels, unsigned m, float as_db, unsigned max_frames, sdr_pfb **out);
int sdr_pfb_destroy(sdr_pfb *pfb);
int sdr_pfb_reset(sdr_pfb *pfb);                                /* firpfbch_crcf_reset */
int sdr_pfb_execute(sdr_pfb *pfb, const sdr_cf32 *iq, unsigned n, sdr_cf32 *out);              /* host buffers */
int sdr_pfb_execute_device(sdr_pfb *pfb, const void *d_iq, unsigned n, void *d_out);           /* device buffers, engine stream */

/* ---- audio-side consumers kept on the device ------------------------------------------------------
 * Mixer = the RtAudio callback of the reference, audioCallback (src/audio/AudioThread.cpp:88-243): every bound source is one
 * AudioThread (inputQueue of per-block AudioThreadInput entries, currentInput, audioQueuePtr, gain, active -- all in HBM).
 * sdr_mixer_bind = AudioThread::bindThread for a demodulator of this engine: after every block its audio, peak and channel count
 * join the source's queue without leaving the GPU.  sdr_mixer_render(n) = one callback for n frames: interleaved stereo, sources
 * summed in binding order, the peak normalisation of :231-238.  sdr_mixer_push queues an AudioThreadInput from the host. */
typedef struct sdr_mixer sdr_mixer;
int sdr_mixer_create(sdr_engine *eng, int sample_rate, unsigned max_frames, sdr_mixer **out);
int sdr_mixer_destroy(sdr_mixer *m);
int sdr_mixer_bind(sdr_mixer *m, sdr_chan *chan, float gain, int *source_index);
int sdr_mixer_add_source(sdr_mixer *m, float gain, int *source_index);
int sdr_mixer_set_gain(sdr_mixer *m, int source, float gain);                  /* AudioThread::setGain (:595-604): clamped to [0, 2] */
/* AudioThread::setActive (:557-589): activation binds the source to the controller again -- at the END of the mix order;
 * deactivation removes it; every call flushes the source's queue (currentInput and audioQueuePtr stay).            */
int sdr_mixer_set_active(sdr_mixer *m, int source, int active);
int sdr_mixer_push(sdr_mixer *m, int source, const float *data, unsigned n, int channels, float peak, int sample_rate);
int sdr_mixer_render(sdr_mixer *m, unsigned n_frames, float *out_stereo /* 2 n_frames */);
/* AudioFileWAV::writePayloadToFileStream (src/audio/AudioFileWAV.cpp:143-165): the channel's last block as 16-bit PCM words */
int sdr_chan_fetch_pcm16(sdr_chan *chan, short *out, unsigned cap, unsigned *n);
/* ScopeVisualProcessor (src/process/ScopeVisualProcessor.cpp:14-217) on a channel's audio: waveform points {x, y} and the audio
 * spectrum's points {x, y} with its running averages; floor_ceil = {fft_floor_maa, fft_ceil_maa} */
typedef struct sdr_scope sdr_scope;
int sdr_scope_create(sdr_engine *eng, sdr_chan *chan, unsigned fft_size, unsigned max_samples, sdr_scope **out);
int sdr_scope_destroy(sdr_scope *s);
int sdr_scope_process(sdr_scope *s, float *wave_xy, unsigned *n_wave, float *spec_xy, unsigned *n_spec, double *floor_ceil);

/* ---- design introspection (host only; used by the tests to pin filter tables against the reference) ---------------- */
int sdr_design_kaiser(unsigned n, float fc, float as_db, float mu, float *h);
int sdr_design_nco_table(float *sintab1024);
uint32_t sdr_design_nco_step(float dtheta);
/* plan of msresamp_{crcf,rrrf}_create(rate, as_db): outputs may be NULL.  pfb: [256*14] floats.  m: [16] ints. */
int sdr_design_resampler(float rate, float as_db, int *interp, int *num_halfband, float *rate_arb, uint32_t *arb_step,
                         float *pfb, int *m);
int sdr_design_halfband_branch(int m, float as_db, float *h1 /* 2m */);
int sdr_design_am_dc_block(float *h51);
int sdr_design_hilbert_q(float *hq10);
int sdr_design_butter6_sos(float *sos18 /* 3 x {b0,b1,b2,a0,a1,a2} */);
/* iirfilt_crcf_create_prototype(CHEBY2, BANDPASS, SOS, 5, fc, f0, 1, 60), ModemFMStereo.cpp:136: 5 x {b0,b1,b2,a0,a1,a2} */
int sdr_design_cheby2_bp5_sos(float fc, float f0, float *sos30);
int sdr_design_hilbert_q_as(float as_db, float *hq10);         /* firhilbf_create(5, as): ModemFMStereo.cpp:138, ModemCW.cpp:22 */

/* ---- group: several GPUs behind the same boundary -------------------------------------------------------
 * The reference hands every block to every demodulator (src/sdr/SDRPostThread.cpp:294-298); channels are independent given
 * the block.  A group places channel c (in creation order) on member c mod N, keeps spectra on member 0 and moves the raw
 * block to every member with ONE ncclBroadcast over NVLink per block; every member runs its own (bit-identical) DC blocker
 * and returns its own channels' results -- sdr_chan_* / sdr_spectrum_* calls work on a group's handles unchanged.
 * sdr_group_create: one process drives n GPUs (CubicSDR is one process, src/CubicSDR.cpp:340-397); devices[0] takes the
 * host block.  sdr_group_create_rank: one process per GPU (torch.distributed.run); nccl_id = the 128 bytes rank 0 got from
 * sdr_group_unique_id, handed to the other ranks by the launcher; NULL = the caller moves the block itself and passes it
 * to sdr_group_submit on every rank.  In that form every rank makes the same sdr_group_chan_create /
 * sdr_group_spectrum_create calls in the same order and receives *out = NULL for objects that live on another rank.      */
typedef struct sdr_group sdr_group;
int sdr_group_create(const int *devices, int n, unsigned max_block, sdr_group **out);
int sdr_group_create_rank(int device, int rank, int world, const void *nccl_id, unsigned max_block, sdr_group **out);
int sdr_group_unique_id(void *id128);
int sdr_group_destroy(sdr_group *g);
/* n_local members in this process, total members of the group, transport (0 none, 1 NCCL broadcast, 2 peer copies) */
int sdr_group_info(sdr_group *g, int *n_local, int *total, int *transport);
int sdr_group_engine(sdr_group *g, int i, sdr_engine **eng);
int sdr_group_chan_create(sdr_group *g, const char *modem, long long sample_rate, long long bandwidth, int audio_rate,
                          sdr_chan **out);
int sdr_group_spectrum_create(sdr_group *g, unsigned fft_size, unsigned lines_per_second, sdr_spectrum **out);
/* One block for the whole group (SDRPostThread::run, src/sdr/SDRPostThread.cpp:178-187): iq is host memory on the root
 * (rank 0; pinned memory is sent by DMA from where it lies, a device pointer is accepted for resident blocks).  Enqueues only: DMA to GPU 0, the broadcast, one pass of the
 * path per member.  sdr_group_wait / sdr_group_wait_block are sdr_engine_wait / sdr_engine_wait_block on every member.    */
int sdr_group_submit(sdr_group *g, const sdr_cf32 *iq, unsigned n, long long center_freq, long long sample_rate);
int sdr_group_wait(sdr_group *g);
int sdr_group_wait_block(sdr_group *g, int lag);
/* 0: leave results in device memory (throughput measurements with resident blocks); default 1 */
int sdr_group_set_host_results(sdr_group *g, int on);

#ifdef __cplusplus
}
#endif
#endif
