{-# LANGUAGE ForeignFunctionInterface #-}
-- | The one function of pure-zlib's "Codec.Compression.Zlib.Deflate" that anything outside the library uses
-- (its test-suite, test/Test.hs:107-120): canonical Huffman code assignment (reference: Deflate.hs:261-288),
-- here computed by the device table builder through pz_compute_code_values.  UNBUILT (no GHC in the image).
module Codec.Compression.Zlib.Deflate (computeCodeValues) where

import Data.Int (Int32)
import Foreign
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

foreign import ccall safe "pz_compute_code_values"
  c_pz_compute_code_values :: Ptr Int32 -> Ptr Int32 -> CInt -> Ptr Int32 -> IO CInt

-- | (symbol, code length) pairs in, (symbol, code length, code) triples out, ascending by symbol; symbols of
-- length 0 are dropped, as in the reference.
computeCodeValues :: [(Int, Int)] -> [(Int, Int, Int)]
computeCodeValues pairs = unsafePerformIO $
  withArrayLen (map (fromIntegral . fst) pairs) $ \n psym ->
    withArray (map (fromIntegral . snd) pairs) $ \plen ->
      allocaArray (3 * max n 1) $ \pout -> do
        m <- c_pz_compute_code_values psym plen (fromIntegral n) pout
        flat <- peekArray (3 * fromIntegral (max m 0)) pout
        return (triples (map fromIntegral flat))
 where
  triples (a : b : c : rest) = (a, b, c) : triples rest
  triples _ = []
