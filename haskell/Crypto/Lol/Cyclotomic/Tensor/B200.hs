{-# LANGUAGE DataKinds, FlexibleContexts, KindSignatures, ScopedTypeVariables, TypeFamilies #-}
-- | The tensor tag 'GT': Lol's @Tensor@ methods on the B200 backend (libalchemy_b200.so), the drop-in for the tag @CT@ of
-- @Crypto.Lol.Cyclotomic.Tensor.CPP@ that the reference selects at examples/Arithmetic.hs:19,23, examples/HomomRLWR.hs:24,47 and
-- examples/Tunnel.hs:20,41 (cabal dependency @lol-cpp@, alchemy.cabal:74,93,117).  Nothing under Crypto/Alchemy names a backend.
--
-- STATUS: written against the surveyor's recollection of the @lol@ 0.7 class surface ([RECALL]; the package is not in /root/reference)
-- and NOT COMPILED -- the build container has no GHC (SURVEY.md F3).  The raw bindings in "Crypto.Lol.Cyclotomic.Tensor.B200.FFI" are
-- generated from the ABI table and checked by the test suite; this module shows how the class methods map onto them.  The same
-- mapping runs, tested, in Python: alchemy_b200/tensor.py (class GT) and alchemy_b200/symm.py.
--
-- Representation: the same as lol-cpp's CT -- a Storable vector of phi(m) elements, RNS tuples interleaved -- so that every
-- instance lol-cpp derives generically (Eq, Show, Random, NFData, fmapT, zipWithT ...) carries over unchanged.  Each method below is
-- ONE call of the host-buffer entry 'c_abTensorOpHost' (interleaved layout, in place on a copy), which is exactly what a
-- @foreign import ccall@ of lol-cpp does today.  Keeping a chain of ciphertext operations on the device -- where the throughput
-- is -- needs the batched device entries (ab_*_dev, ab_mul_key_switch_dev, ab_tunnel_dev); see INTEGRATION.md section 3.
module Crypto.Lol.Cyclotomic.Tensor.B200 (GT) where

import Control.Monad (when)
import Data.IORef
import Data.Int (Int32, Int64)
import qualified Data.Map.Strict as M
import qualified Data.Vector.Storable as SV
import qualified Data.Vector.Storable.Mutable as SM
import Foreign.C.String (peekCString)
import Foreign.ForeignPtr
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (withArrayLen)
import Foreign.Ptr
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import Crypto.Lol.Cyclotomic.Tensor.B200.FFI
-- import Crypto.Lol.Cyclotomic.Tensor (Tensor (..))   -- [RECALL] class and helper names below follow lol-0.7
-- import Crypto.Lol.Factored; import Crypto.Lol.Types.ZqBasic; import Crypto.Lol.CRTrans

-- | tensor over the m-th cyclotomic with coefficients r, as a Storable vector (lol-cpp: @newtype CT m r = CT' (Vector r)@)
newtype GT (m :: k) r = GT (SV.Vector r)

-- op codes of ab_tensor_op_host (include/alchemy_b200.h)
data Op = OpCRT | OpCRTInv | OpL | OpLInv | OpMulGPow | OpMulGDec | OpDivGPow | OpDivGDec | OpMulGCRT | OpDivGCRT deriving (Enum)

layoutInterleaved :: Int32
layoutInterleaved = 1

-- | one plan per (prime powers of m, moduli): device tables are built once and shared by every call (plans are immutable and re-entrant)
type PlanKey = ([(Int32, Int32)], [Int64], Bool)
{-# NOINLINE planCache #-}
planCache :: IORef (M.Map PlanKey (ForeignPtr ()))
planCache = unsafePerformIO (newIORef M.empty)

check :: Int32 -> IO ()
check 0 = return ()
check rc = c_abLastError >>= peekCString >>= \msg -> ioError (userError ("alchemy_b200 error " ++ show rc ++ ": " ++ msg))

-- | @withPlan pps moduli crt@: pps = prime-power factorisation of m (Lol @ppsFact@), moduli = the RNS moduli of the element type
-- (one per component of a @ZqBasic@ tuple), crt = whether CRT tables are needed (False: L, G in Pow/Dec over any modulus)
withPlan :: [(Int32, Int32)] -> [Int64] -> Bool -> (Ptr () -> IO a) -> IO a
withPlan pps qs crt act = do
  cache <- readIORef planCache
  fp <- case M.lookup (pps, qs, crt) cache of
    Just fp -> return fp
    Nothing -> alloca $ \out ->
      withArrayLen [PrimeExp p e | (p, e) <- pps] $ \npe pe ->
        withArrayLen qs $ \nq pq -> do
          check =<< if crt then c_abPlanCreate out pe (fromIntegral npe) pq (fromIntegral nq) nullPtr   -- NULL roots: principalRootUnity
                           else c_abPlanCreateNocrt out pe (fromIntegral npe) pq (fromIntegral nq)
          fp <- newForeignPtr p_abPlanDestroy =<< peek out
          atomicModifyIORef' planCache (\c -> (M.insert (pps, qs, crt) fp c, ()))
          return fp
  withForeignPtr fp act

-- | a Tensor method over Z_q tuples: copy, run in place on the device through the host-buffer entry, freeze
tensorOp :: SV.Storable r => Op -> [(Int32, Int32)] -> [Int64] -> SV.Vector r -> SV.Vector r
tensorOp op pps qs v = unsafePerformIO $ do
  mv <- SV.thaw v
  SM.unsafeWith mv $ \p ->
    withPlan pps qs (fromEnum op <= fromEnum OpCRTInv || fromEnum op >= fromEnum OpMulGCRT) $ \plan ->
      check =<< c_abTensorOpHost plan (fromIntegral (fromEnum op)) (castPtr p) 1 layoutInterleaved
  SV.unsafeFreeze mv

{- The instance, method by method ([RECALL] signatures of lol-0.7's class Tensor; `pps` = ppsFact @m, `qs` = moduli of r):

instance Tensor GT where
  type TElt GT r = (Storable r, TensorRq r)          -- TensorRq: reflects the tuple's moduli, as lol-cpp's Dispatch class does

  l       (GT v) = GT (tensorOp OpL       pps qs v)  -- tensorLRq      Dec -> Pow
  lInv    (GT v) = GT (tensorOp OpLInv    pps qs v)  -- tensorLInvRq
  mulGPow (GT v) = GT (tensorOp OpMulGPow pps qs v)  -- tensorGPowRq
  mulGDec (GT v) = GT (tensorOp OpMulGDec pps qs v)  -- tensorGDecRq
  divGPow (GT v) = Just (GT (tensorOp OpDivGPow pps qs v))   -- over Z_q the division always succeeds; over Z the generic Haskell path stays
  divGDec (GT v) = Just (GT (tensorOp OpDivGDec pps qs v))

  crtFuncs = do                                       -- exists iff every modulus has an m-th root of unity (AB_ERR_NO_ROOT otherwise)
    _ <- crtInfo @m                                   -- the CRTrans monad decides existence exactly as for CT
    return ( scalarCRT                                -- replicate
           , \(GT v) -> GT (tensorOp OpMulGCRT pps qs v)     -- tensorGCRTRq
           , \(GT v) -> GT (tensorOp OpDivGCRT pps qs v)     -- tensorGInvCRTRq
           , \(GT v) -> GT (tensorOp OpCRT     pps qs v)     -- tensorCRTRq
           , \(GT v) -> GT (tensorOp OpCRTInv  pps qs v) )   -- tensorCRTInvRq (includes the mhat^-1 scaling)

  -- twacePowDec / embedPow / embedDec / coeffs / powBasisPow and crtExtFuncs (twaceCRT, embedCRT): index tables live in an ab_ext
  -- (c_abExtCreate lo hi); one call of c_abTwacePowDecDev / c_abEmbedPowDev / c_abEmbedDecDev / c_abTwaceCrtDev / c_abEmbedCrtDev /
  -- c_abCoeffsDev on device copies.  lol-cpp does these in Haskell with `backpermute`; that code keeps working on GT unchanged
  -- because the representation is the same vector, so they are an optimisation, not a requirement of the swap.
  -- zipWithT (*) on Z_q tuples = c_abMulHost (mulRq); fmapT / zipWithT / unzipT / scalarPow / tGaussianDec / gSqNormDec: generic
  -- Storable-vector code shared with lol-cpp.
-}

-- | swap at the three sites of the reference:
--
-- > import Crypto.Lol.Cyclotomic.Tensor.B200            -- was: Crypto.Lol.Cyclotomic.Tensor.CPP
-- > type PT = PNoiseCyc PNZ GT F4 (Zq 7)                -- was: ... CT ...      (examples/Arithmetic.hs:23)
_unused :: ()
_unused = when False (return ()) `seq` ()
