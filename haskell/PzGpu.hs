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
-- is exercised through the identical C symbols by tests/ (ctypes) and host/examples_gpu.cpp.
module PzGpu
  ( ModelDesc, Posterior (..), rw1d, stt, sttDelayed, mttTrack
  , inferGPU, zparticlesGPU, runStreamGPU
  ) where

import Control.Monad (when)
import Data.Word (Word64)
import Data.Int (Int32, Int64)
import Foreign
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types

import Util.ZStream (ZStream (..))   -- haskell/src/Util/ZStream.hs:16
import qualified Util.ZStream as ZS

-- pz_model_desc is 8*4 + 8*(36+36+18+9+6) + 72 = 944 bytes of plain old data.
newtype ModelDesc = ModelDesc (ForeignPtr ())
data PzFilter

sizeofModelDesc, sizeofSummary :: Int
sizeofModelDesc = 944
sizeofSummary = 144

data Posterior = Posterior
  { postStep :: Int64
  , postLogEvidenceInc :: Double
  , postEss :: Double
  , postMean :: [Double]     -- ^ posterior mean of the state (weighted particle average; cf.
                             --   `average particles`, Examples/Demo.hs:207)
  , postVar :: [Double]
  , postParticles :: Int -> IO [[Double]]  -- ^ every k-th equally weighted particle after
                             --   resampling: the `[b]` of zparticles' (Inference.hs:104)
  }

foreign import ccall unsafe "pz_model_rw1d" c_model_rw1d :: Ptr () -> CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "pz_model_stt" c_model_stt :: Ptr () -> IO CInt
foreign import ccall unsafe "pz_model_mtt_track" c_model_mtt_track :: Ptr () -> IO CInt
foreign import ccall safe "pz_filter_create"
  c_filter_create :: Ptr () -> Word64 -> Word64 -> CInt -> Ptr (Ptr PzFilter) -> IO CInt
foreign import ccall safe "pz_filter_step"
  c_filter_step :: Ptr PzFilter -> Ptr CDouble -> Int32 -> Int32 -> Ptr () -> IO CInt
foreign import ccall safe "pz_filter_read_particles"
  c_read_particles :: Ptr PzFilter -> Int32 -> Ptr CDouble -> Int64 -> IO CInt
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

-- | infer: model x particle count -> stream of posteriors.  One `step` is one call of
-- pz_filter_step (the whole of zparticles', Inference.hs:101-104, on the GPU).
inferGPU :: ModelDesc -> Int -> Word64 -> IO (ZStream IO [Double] Posterior)
inferGPU (ModelDesc dfp) n seed = do
  h <- withForeignPtr dfp $ \dp -> alloca $ \out -> do
    c_filter_create dp (fromIntegral n) seed 0 out >>= check
    peek out
  fh <- newForeignPtr p_filter_destroy h
  nx <- pure 6  -- mean/var arrays are PZ_MAX_NX wide; callers take the first nx entries
  let stepf () obs = withForeignPtr fh $ \f ->
        withArrayLen (map realToFrac obs) $ \m po ->
        allocaBytes sizeofSummary $ \ps -> do
          c_filter_step f po 1 (fromIntegral m) ps >>= check
          st <- peekByteOff ps 0 :: IO Int64
          le <- peekByteOff ps 8 :: IO CDouble
          es <- peekByteOff ps 16 :: IO CDouble
          mu <- peekArray nx (ps `plusPtr` 24) :: IO [CDouble]
          va <- peekArray nx (ps `plusPtr` 72) :: IO [CDouble]
          let particles k = withForeignPtr fh $ \f' -> do
                nl <- c_n_local f'
                let rows = (fromIntegral nl + k - 1) `div` k
                allocaArray (rows * nx) $ \buf -> do
                  c_read_particles f' (fromIntegral k) buf (fromIntegral (rows * nx)) >>= check
                  flat <- peekArray (rows * nx) buf
                  pure (chunks nx (map realToFrac flat))
          pure ((), Posterior st (realToFrac le) (realToFrac es) (map realToFrac mu) (map realToFrac va) particles)
  pure (ZS.fromStep stepf ())          -- Util/ZStream.hs:20-21
  where
    chunks _ [] = []
    chunks k xs = let (a, b) = splitAt k xs in a : chunks k b

-- | Drop-in for `zparticles n model` when `model` is one of the supported families.
zparticlesGPU :: Int -> ModelDesc -> IO (ZStream IO [Double] Posterior)
zparticlesGPU n d = inferGPU d n 0

-- | ZS.runStream (Util/ZStream.hs:65-69) over a finite observation list.
runStreamGPU :: (Posterior -> IO ()) -> ZStream IO [Double] Posterior -> [[Double]] -> IO ()
runStreamGPU _ _ [] = pure ()
runStreamGPU act (ZStream f) (y : ys) = do
  (s', post) <- f y
  act post
  runStreamGPU act s' ys
