{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE Arrows #-}
-- | PzGpu: the reference-side binding of libpzgpu.so (include/pzgpu.h).
--
-- This is the file a maintainer of psg-mit/probzelus-haskell adds under haskell/src/ to
-- route the streaming particle filter through the B200 kernels.  It keeps the shape of
-- the reference API: a model and a particle count give a stream of posteriors,
--
-- >   zparticles   :: MonadSample m => Int -> ZStream (Weighted m) a b -> ZStream m a [b]   -- Inference.hs:107-108
-- >   inferGPU     ::                  ModelDesc -> Int -> Word64 -> IO (ZStream IO [Double] Posterior)
--
-- with the Haskell closure replaced by a kernel-side descriptor.  NOT COMPILED in this
-- repository's CI: the build image has no GHC (SURVEY.md section 0.9).  Every foreign import below
-- is exercised through the identical C symbols, in the same order and with the same buffer
-- arithmetic, by host/shim_mirror.c (tests/test_host_cli.py), by tests/ (ctypes) and by
-- host/examples_gpu.cpp.
module PzGpu
  ( ModelDesc, Posterior (..), Track (..), CamTrack (..)
  , rw1d, stt, sttDelayed, mttTrack, mtt, multiCamera, walker, betaBernoulli, withDouble, withDevices
  , inferGPU, inferGPUOn, zparticlesGPU, zdsparticlesGPU, runStreamGPU, runAllGPU
  , mttGPU, multiCameraGPU
  ) where

import Control.Monad (when, forM)
import Data.Word (Word64)
import Data.Int (Int32, Int64)
import Foreign
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types

import Util.ZStream (ZStream (..))   -- haskell/src/Util/ZStream.hs:16
import qualified Util.ZStream as ZS

-- pz_model_desc (ABI 2) is 1152 bytes of plain old data; the fields this module touches:
--   kind :: Int32 @0, nx :: Int32 @4, ny :: Int32 @8, delayed :: Int32 @20, state_f64 :: Int32 @24.
newtype ModelDesc = ModelDesc (ForeignPtr ())
data PzFilter

sizeofModelDesc, sizeofSummary :: Int
sizeofModelDesc = 1152
sizeofSummary = 144

data Posterior = Posterior
  { postStep :: Int64
  , postLogEvidenceInc :: Double
  , postEss :: Double
  , postMean :: [Double]     -- ^ posterior mean of the state, nx entries (weighted particle average; cf.
                             --   `average particles`, Examples/Demo.hs:207)
  , postVar :: [Double]
  , postParticles :: Int -> IO [[Double]]  -- ^ every k-th equally weighted particle after
                             --   resampling: the `[b]` of zparticles' (Inference.hs:104)
  }

-- | One entry of the tracker's `Map id (mean, cov)` (MultiTargetTracking.hs:133-139).
data Track = Track { trackId :: Int32, trackMean :: [Double], trackCov :: [[Double]] }

-- | `MarginalTrack` of Examples/MultiCamera.hs:48-69: the box marginal (its covariance is diagonal), trackID, identity, camera.
data CamTrack = CamTrack { camTrackId :: Int32, camIdentity :: Int32, camCamera :: Int32, camStart :: Double
                         , camBoxMean :: [Double], camBoxVar :: [Double] }

foreign import ccall unsafe "pz_model_rw1d" c_model_rw1d :: Ptr () -> CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "pz_model_stt" c_model_stt :: Ptr () -> IO CInt
foreign import ccall unsafe "pz_model_mtt_track" c_model_mtt_track :: Ptr () -> IO CInt
foreign import ccall unsafe "pz_model_mtt" c_model_mtt :: Ptr () -> IO CInt
foreign import ccall unsafe "pz_model_multicam" c_model_multicam :: Ptr () -> IO CInt
foreign import ccall unsafe "pz_filter_read_mcam_tracks"
  c_read_mcam_tracks :: Ptr PzFilter -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr CDouble -> Ptr CDouble
                     -> Ptr CDouble -> Ptr Int32 -> Ptr CDouble -> Ptr CDouble -> Int64 -> IO CInt
foreign import ccall unsafe "pz_model_walker" c_model_walker :: Ptr () -> IO CInt
foreign import ccall unsafe "pz_model_beta_bernoulli" c_model_beta_bernoulli :: Ptr () -> CDouble -> CDouble -> IO CInt
foreign import ccall safe "pz_filter_create"
  c_filter_create :: Ptr () -> Word64 -> Word64 -> CInt -> Ptr (Ptr PzFilter) -> IO CInt
foreign import ccall safe "pz_filter_create_multi"
  c_filter_create_multi :: Ptr () -> Word64 -> Word64 -> CInt -> Ptr CInt -> Ptr (Ptr PzFilter) -> IO CInt
foreign import ccall safe "pz_filter_step"
  c_filter_step :: Ptr PzFilter -> Ptr CDouble -> Int32 -> Int32 -> Ptr () -> IO CInt
foreign import ccall safe "pz_filter_run"
  c_filter_run :: Ptr PzFilter -> Ptr CDouble -> Int64 -> Int32 -> Ptr () -> IO CInt
foreign import ccall safe "pz_filter_read_particles"
  c_read_particles :: Ptr PzFilter -> Int32 -> Ptr CDouble -> Int64 -> IO CInt
foreign import ccall safe "pz_filter_read_tracks"
  c_read_tracks :: Ptr PzFilter -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr CDouble -> Ptr CDouble -> Int64 -> IO CInt
foreign import ccall safe "pz_filter_read_posterior"
  c_read_posterior :: Ptr PzFilter -> Ptr CDouble -> Ptr CDouble -> IO CInt  -- MDistr (mean, cov) of a delayed filter
foreign import ccall unsafe "pz_filter_n_local" c_n_local :: Ptr PzFilter -> IO Int64
foreign import ccall unsafe "&pz_filter_destroy" p_filter_destroy :: FunPtr (Ptr PzFilter -> IO ())
foreign import ccall unsafe "pz_last_error" c_last_error :: IO CString

check :: CInt -> IO ()
check rc = when (rc /= 0) $ do
  msg <- c_last_error >>= peekCString
  ioError (userError ("libpzgpu: " ++ show rc ++ ": " ++ msg))

mkDesc :: (Ptr () -> IO CInt) -> IO ModelDesc
mkDesc fill = do
  fp <- mallocForeignPtrBytes sizeofModelDesc
  withForeignPtr fp (\p -> fill p >>= check)
  pure (ModelDesc fp)

-- | kalman1DObserved (Examples/Demo.hs:184-198): x' ~ N(x, 1), y ~ N(x', 3), doubled score.
rw1d :: Double -> Double -> IO ModelDesc
rw1d q r = mkDesc (\p -> c_model_rw1d p (realToFrac q) (realToFrac r))

-- | Examples/SingleTargetTracking.hs:28-54 with delay = False (the bootstrap filter).
stt :: IO ModelDesc
stt = mkDesc c_model_stt

-- | The same model with delay = True (SingleTargetTracking.hs:36-38): the state stays a symbolic
-- Gaussian marginal; every particle carries the exact Kalman posterior (pz_model_desc.delayed,
-- a 32-bit field at byte offset 20).
sttDelayed :: IO ModelDesc
sttDelayed = do
  ModelDesc fp <- stt
  withForeignPtr fp (\p -> pokeByteOff p 20 (1 :: Int32))
  pure (ModelDesc fp)

-- | The per-track model of Examples/MultiTargetTracking.hs:56-73.
mttTrack :: IO ModelDesc
mttTrack = mkDesc c_model_mtt_track

-- | The two-camera tracker, Examples/MultiCamera.hs:40-227.
multiCamera :: IO ModelDesc
multiCamera = mkDesc c_model_multicam

-- | The multi-target tracker, Examples/MultiTargetTracking.hs:26-145 (`examples mtt N`).
mtt :: IO ModelDesc
mtt = mkDesc c_model_mtt

-- | Examples/Walker.hs under `particles N` (cumulative weights).
walker :: IO ModelDesc
walker = mkDesc c_model_walker

-- | betaBernoulli a b (Examples/DelayedSamplingAbstract.hs:57-69) under delayed sampling.
betaBernoulli :: Double -> Double -> IO ModelDesc
betaBernoulli a b = mkDesc (\p -> c_model_beta_bernoulli p (realToFrac a) (realToFrac b))

-- | State and arithmetic in Double, the reference's own precision (pz_model_desc.state_f64 at byte offset 24).
withDouble :: ModelDesc -> IO ModelDesc
withDouble (ModelDesc fp) = withForeignPtr fp (\p -> pokeByteOff p 24 (1 :: Int32)) >> pure (ModelDesc fp)

descNx :: ModelDesc -> IO Int
descNx (ModelDesc fp) = withForeignPtr fp $ \p -> fromIntegral <$> (peekByteOff p 4 :: IO Int32)

readSummary :: Int -> Ptr () -> IO (Int64, Double, Double, [Double], [Double])
readSummary nx ps = do
  st <- peekByteOff ps 0 :: IO Int64
  le <- peekByteOff ps 8 :: IO CDouble
  es <- peekByteOff ps 16 :: IO CDouble
  mu <- peekArray nx (ps `plusPtr` 24) :: IO [CDouble]
  va <- peekArray nx (ps `plusPtr` 72) :: IO [CDouble]
  pure (st, realToFrac le, realToFrac es, map realToFrac mu, map realToFrac va)

chunks :: Int -> [a] -> [[a]]
chunks _ [] = []
chunks k xs = let (a, b) = splitAt k xs in a : chunks k b

-- | One process, several devices (SURVEY 8b: single-threaded host, the filter owns every device it is given).
withDevices :: [Int] -> ModelDesc -> Int -> Word64 -> IO (ZStream IO [Double] Posterior)
withDevices devs d n seed = inferWith (\dp out -> withArrayLen (map fromIntegral devs) $ \k pd ->
                                        c_filter_create_multi dp (fromIntegral n) seed (fromIntegral k) pd out) d

-- | infer: model x particle count -> stream of posteriors.  One `step` is one call of
-- pz_filter_step (the whole of zparticles', Inference.hs:101-104, on the GPU).
inferGPU :: ModelDesc -> Int -> Word64 -> IO (ZStream IO [Double] Posterior)
inferGPU d n seed = inferGPUOn 0 d n seed

inferGPUOn :: Int -> ModelDesc -> Int -> Word64 -> IO (ZStream IO [Double] Posterior)
inferGPUOn dev d n seed = inferWith (\dp out -> c_filter_create dp (fromIntegral n) seed (fromIntegral dev) out) d

inferWith :: (Ptr () -> Ptr (Ptr PzFilter) -> IO CInt) -> ModelDesc -> IO (ZStream IO [Double] Posterior)
inferWith create d@(ModelDesc dfp) = do
  h <- withForeignPtr dfp $ \dp -> alloca $ \out -> do
    create dp out >>= check
    peek out
  fh <- newForeignPtr p_filter_destroy h
  nx <- descNx d   -- rows of pz_filter_read_particles are nx wide; mean / var hold nx entries
  let stepf () obs = withForeignPtr fh $ \f ->
        withArrayLen (map realToFrac obs) $ \m po ->
        allocaBytes sizeofSummary $ \ps -> do
          c_filter_step f po 1 (fromIntegral m) ps >>= check
          (st, le, es, mu, va) <- readSummary nx ps
          let particles k = withForeignPtr fh $ \f' -> do
                nl <- c_n_local f'
                let rows = (fromIntegral nl + k - 1) `div` k
                allocaArray (rows * nx) $ \buf -> do
                  c_read_particles f' (fromIntegral k) buf (fromIntegral (rows * nx)) >>= check
                  flat <- peekArray (rows * nx) buf
                  pure (chunks nx (map realToFrac flat))
          pure ((), Posterior st le es mu va particles)
  pure (ZS.fromStep stepf ())          -- Util/ZStream.hs:20-21

-- | Drop-in for `zparticles n model` when `model` is one of the supported families.
zparticlesGPU :: Int -> ModelDesc -> IO (ZStream IO [Double] Posterior)
zparticlesGPU n d = inferGPU d n 0

-- | Drop-in for `zdsparticles n model` (Inference.hs:113-114): a descriptor with delayed = 1.
zdsparticlesGPU :: Int -> ModelDesc -> IO (ZStream IO [Double] Posterior)
zdsparticlesGPU = zparticlesGPU

-- | ZS.runStream (Util/ZStream.hs:65-69) over a finite observation list.
runStreamGPU :: (Posterior -> IO ()) -> ZStream IO [Double] Posterior -> [[Double]] -> IO ()
runStreamGPU _ _ [] = pure ()
runStreamGPU act (ZStream f) (y : ys) = do
  (s', post) <- f y
  act post
  runStreamGPU act s' ys

-- | The whole observation list in one call (pz_filter_run): the steps run back to back on the device and the
-- summaries come back together.  Returns (step, log-evidence increment, ESS, mean, var) per step.
runAllGPU :: ModelDesc -> Int -> Word64 -> [[Double]] -> IO [(Int64, Double, Double, [Double], [Double])]
runAllGPU d@(ModelDesc dfp) n seed obs = do
  h <- withForeignPtr dfp $ \dp -> alloca $ \out -> c_filter_create dp (fromIntegral n) seed 0 out >>= check >> peek out
  fh <- newForeignPtr p_filter_destroy h
  nx <- descNx d
  let t = length obs
      ny = if null obs then 0 else length (head obs)
  withForeignPtr fh $ \f -> withArray (map realToFrac (concat obs)) $ \po ->
    allocaBytes (t * sizeofSummary) $ \ps -> do
      c_filter_run f po (fromIntegral t) (fromIntegral ny) ps >>= check
      forM [0 .. t - 1] $ \i -> readSummary nx (ps `plusPtr` (i * sizeofSummary))

-- | `examples mtt N` (test/Examples.hs:23-24, MultiTargetTracking.hs:133-145): each step takes the detections of
-- the frame and returns the posterior summary and, for every k-th resampled particle, its track map.
mttGPU :: Int -> Word64 -> IO (ZStream IO [[Double]] (Posterior, Int -> IO [[Track]]))
mttGPU n seed = do
  d@(ModelDesc dfp) <- mtt
  h <- withForeignPtr dfp $ \dp -> alloca $ \out -> c_filter_create dp (fromIntegral n) seed 0 out >>= check >> peek out
  fh <- newForeignPtr p_filter_destroy h
  nx <- descNx d
  let stepf () dets = withForeignPtr fh $ \f ->
        withArray (map realToFrac (concat dets)) $ \po ->
        allocaBytes sizeofSummary $ \ps -> do
          c_filter_step f po (fromIntegral (length dets)) 2 ps >>= check
          (st, le, es, mu, va) <- readSummary nx ps
          let tracks k = withForeignPtr fh $ \f' -> do
                nl <- c_n_local f'
                let rows = (fromIntegral nl + k - 1) `div` k
                allocaArray rows $ \pc -> allocaArray (rows * 16) $ \pid ->
                  allocaArray (rows * 16 * 4) $ \pm -> allocaArray (rows * 16 * 16) $ \pv -> do
                    c_read_tracks f' (fromIntegral k) pc pid pm pv (fromIntegral rows) >>= check
                    cnt <- peekArray rows pc
                    forM (zip [0 ..] cnt) $ \(i, c) -> forM [0 .. fromIntegral c - 1] $ \j -> do
                      tid <- peekElemOff pid (i * 16 + j)
                      m <- peekArray 4 (pm `advancePtr` ((i * 16 + j) * 4))
                      v <- peekArray 16 (pv `advancePtr` ((i * 16 + j) * 16))
                      pure (Track tid (map realToFrac m) (chunks 4 (map realToFrac v)))
          pure ((), (Posterior st le es mu va (const (pure [])), tracks))
  pure (ZS.fromStep stepf ())

-- | `MultiCamera.runMTTPF` (Examples/MultiCamera.hs:221-225): each step takes the frame's observation rows
-- (camera, pose[10], box[4], appearance reading) and returns the posterior summary and, for every k-th resampled particle,
-- its list of marginal tracks.
multiCameraGPU :: Int -> Word64 -> IO (ZStream IO [[Double]] (Posterior, Int -> IO [[CamTrack]]))
multiCameraGPU n seed = do
  d@(ModelDesc dfp) <- multiCamera
  h <- withForeignPtr dfp $ \dp -> alloca $ \out -> c_filter_create dp (fromIntegral n) seed 0 out >>= check >> peek out
  fh <- newForeignPtr p_filter_destroy h
  nx <- descNx d
  let stepf () rows = withForeignPtr fh $ \f ->
        withArray (map realToFrac (concat rows)) $ \po ->
        allocaBytes sizeofSummary $ \ps -> do
          c_filter_step f po (fromIntegral (length rows)) 16 ps >>= check
          (st, le, es, mu, va) <- readSummary nx ps
          let tracks k = withForeignPtr fh $ \f' -> do
                nl <- c_n_local f'
                let sel = (fromIntegral nl + k - 1) `div` k
                allocaArray sel $ \pc -> allocaArray (sel * 16) $ \pid -> allocaArray (sel * 16) $ \pidn ->
                  allocaArray (sel * 16) $ \pcam -> allocaArray (sel * 16) $ \pst ->
                  allocaArray (sel * 16 * 4) $ \pm -> allocaArray (sel * 16 * 4) $ \pv -> do
                    c_read_mcam_tracks f' (fromIntegral k) pc pid pidn pcam pst pm pv nullPtr nullPtr nullPtr (fromIntegral sel) >>= check
                    cnt <- peekArray sel pc
                    forM (zip [0 ..] cnt) $ \(i, c) -> forM [0 .. fromIntegral c - 1] $ \j -> do
                      tid <- peekElemOff pid (i * 16 + j)
                      idn <- peekElemOff pidn (i * 16 + j)
                      cam <- peekElemOff pcam (i * 16 + j)
                      s0 <- peekElemOff pst (i * 16 + j)
                      m <- peekArray 4 (pm `advancePtr` ((i * 16 + j) * 4))
                      v <- peekArray 4 (pv `advancePtr` ((i * 16 + j) * 4))
                      pure (CamTrack tid idn cam (realToFrac s0) (map realToFrac m) (map realToFrac v))
          pure ((), (Posterior st le es mu va (const (pure [])), tracks))
  pure (ZS.fromStep stepf ())
