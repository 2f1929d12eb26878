{-# LANGUAGE ForeignFunctionInterface #-}
-- | Drop-in for pure-zlib's "Codec.Compression.Zlib" (reference: src/Codec/Compression/Zlib.hs:3-8)
-- on top of libpzcuda.so (include/pzcuda.h).  Same export list, same types, same verdicts; the
-- work happens on a B200.  UNBUILT in this repository (no GHC in the build image): it is the
-- binding a maintainer adds, kept small and mechanical.  See INTEGRATION.md.
module Codec.Compression.Zlib (
  DecompressionError (..),
  ZlibDecoder (NeedMore, Chunk, Done, DecompError),
  decompress,
  decompressIncremental,
  -- * extension: many independent streams per kernel launch
  decompressBatch,
  -- * extension: the same decoder behind gzip members (RFC 1952) and raw deflate (the reference's README TODO)
  Framing (..),
  decompressBatchWith,
  decompressGzip,
  decompressRaw,
  -- * extension: many multi-chunk streams advanced together, one launch per round of chunks
  decompressMany,
) where

import Control.Exception (ErrorCall (..), Exception, throw)
import Control.Monad (forM, when)
import Control.Monad.ST (ST)
import Control.Monad.ST.Unsafe (unsafeIOToST)
import qualified Data.ByteString as S
import qualified Data.ByteString.Internal as SI
import qualified Data.ByteString.Lazy as L
import qualified Data.ByteString.Unsafe as SU
import Data.Int (Int32, Int64)
import Data.Word (Word32, Word64, Word8)
import Foreign
import Foreign.ForeignPtr.Unsafe (unsafeForeignPtrToPtr)
import Foreign.C.String (CString, peekCString, peekCStringLen)
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

-- | Monad.hs:87-104 (constructors, Eq, and the Show prefixes are the reference's).
data DecompressionError
  = HuffmanTreeError String
  | FormatError String
  | DecompressionError String
  | HeaderError String
  | ChecksumError String
  deriving (Eq)

instance Show DecompressionError where
  show x = case x of
    HuffmanTreeError s -> "Huffman tree manipulation error: " ++ s
    FormatError s -> "Block format error: " ++ s
    DecompressionError s -> "Decompression error: " ++ s
    HeaderError s -> "Header error: " ++ s
    ChecksumError s -> "Checksum error: " ++ s

-- | Monad.hs:104.  Callers `throwIO` / `catch` this type (it is exported with its constructors,
-- Zlib.hs:4); `Typeable` is derived automatically by every GHC the reference supports (>= 7.10).
instance Exception DecompressionError

-- | Monad.hs:163-167.
data ZlibDecoder s
  = NeedMore (S.ByteString -> ST s (ZlibDecoder s))
  | Chunk S.ByteString (ST s (ZlibDecoder s))
  | Done
  | DecompError DecompressionError

-- pz_result, 48 bytes (include/pzcuda.h)
data PzResult = PzResult
  { rStatus :: !Int32, rDetail :: !Int32, rOutLen :: !Word64, rAdlerC :: !Word32, rAdlerS :: !Word32
  , rBitPos :: !Word64, rP0 :: !Int64, rP1 :: !Int64 }

instance Storable PzResult where
  sizeOf _ = 48
  alignment _ = 8
  peek p = PzResult <$> peekByteOff p 0 <*> peekByteOff p 4 <*> peekByteOff p 8 <*> peekByteOff p 16
                    <*> peekByteOff p 20 <*> peekByteOff p 24 <*> peekByteOff p 32 <*> peekByteOff p 40
  poke p r = do
    pokeByteOff p 0 (rStatus r); pokeByteOff p 4 (rDetail r); pokeByteOff p 8 (rOutLen r)
    pokeByteOff p 16 (rAdlerC r); pokeByteOff p 20 (rAdlerS r); pokeByteOff p 24 (rBitPos r)
    pokeByteOff p 32 (rP0 r); pokeByteOff p 40 (rP1 r)

data PzStream
data PzOutputs

-- `safe`: the calls block on the GPU.  The library is thread-safe and deterministic, which is
-- what lets `decompress` stay a pure function (unsafePerformIO below).
foreign import ccall safe "pz_inflate_batch"
  c_pz_inflate_batch :: Ptr (Ptr Word8) -> Ptr CSize -> Ptr (Ptr Word8) -> Ptr CSize -> CSize -> Ptr PzResult -> Word32 -> IO CInt
foreign import ccall safe "pz_inflate_sizes"
  c_pz_inflate_sizes :: Ptr (Ptr Word8) -> Ptr CSize -> CSize -> Ptr PzResult -> IO CInt
foreign import ccall safe "pz_decompress_batch"
  c_pz_decompress_batch :: Ptr (Ptr Word8) -> Ptr CSize -> CSize -> Ptr PzResult -> Ptr (Ptr Word8) -> Ptr (Ptr PzOutputs) -> Word32 -> IO CInt
foreign import ccall unsafe "&pz_outputs_free"
  p_pz_outputs_free :: FunPtr (Ptr PzOutputs -> IO ())
foreign import ccall unsafe "pz_strerror"
  c_pz_strerror :: Ptr PzResult -> CString -> CSize -> IO CSize
foreign import ccall unsafe "pz_last_error"
  c_pz_last_error :: IO CString
foreign import ccall safe "pz_stream_new"
  c_pz_stream_new :: IO (Ptr PzStream)
foreign import ccall safe "pz_stream_feed"
  c_pz_stream_feed :: Ptr PzStream -> Ptr Word8 -> CSize -> IO CInt
foreign import ccall safe "pz_stream_next"
  c_pz_stream_next :: Ptr PzStream -> Ptr (Ptr Word8) -> Ptr CSize -> Ptr PzResult -> IO CInt
foreign import ccall unsafe "&pz_stream_free"
  p_pz_stream_free :: FunPtr (Ptr PzStream -> IO ())
foreign import ccall safe "pz_stream_pump"
  c_pz_stream_pump :: Ptr (Ptr PzStream) -> CSize -> IO CInt
-- (pz_stream_feed_many is the batched counterpart of pz_stream_feed -- one copy across the bus for a round of
-- chunks; decompressMany below keeps to one pz_stream_feed per stream for the sake of brevity)

-- | The verdict as the reference's value: Left e, or the impure exception the reference dies with.
verdict :: Ptr PzResult -> PzResult -> IO (Maybe DecompressionError)
verdict p r = case rStatus r of
  0 -> return Nothing
  6 -> message >>= \m -> throw (ErrorCall m) -- PZ_REF_BOTTOM: array / vector bounds error in the reference
  7 -> throw (ErrorCall "pzcuda: output capacity too small (internal sizing bug)")
  s -> do
    m <- message
    let strip pre = drop (length pre) m
    return . Just $ case s of
      1 -> HuffmanTreeError (strip "Huffman tree manipulation error: ")
      2 -> FormatError (strip "Block format error: ")
      3 -> DecompressionError (strip "Decompression error: ")
      4 -> HeaderError (strip "Header error: ")
      _ -> ChecksumError (strip "Checksum error: ")
 where
  message = allocaBytes 512 $ \buf -> do
    n <- c_pz_strerror p buf 512
    peekCStringLen (buf, fromIntegral (min n 511))

-- | Zlib.hs:32-51 applied to every element, in ONE library call (pz_decompress_batch: sizing launch, output
-- allocation and decode launch inside the library).  The decoded bytes of all streams live in one pinned block
-- owned by the library; the result ByteStrings are slices of it that share one ForeignPtr, whose finalizer
-- (pz_outputs_free) runs when the last of them is collected: no byte is copied on the host on the way out.
-- A lazy ByteString of several chunks whose stream ends before the last chunk is the
-- reference's "Finished with data remaining." (Zlib.hs:48-49); the shim checks that on the host.
decompressBatch :: [L.ByteString] -> [Either DecompressionError L.ByteString]
decompressBatch = decompressBatchWith Zlib

-- | PZ_F_GZIP / PZ_F_RAW of include/pzcuda.h.  With 'Gzip' the words in a 'ChecksumError' are CRC-32s (or ISIZE and the
-- decoded length for "length mismatch"); 'RawDeflate' has no trailer and therefore no checksum verdict.
data Framing = Zlib | Gzip | RawDeflate deriving (Eq, Show)

framingFlags :: Framing -> Word32
framingFlags Zlib = 0
framingFlags Gzip = 0x20
framingFlags RawDeflate = 0x40

decompressGzip, decompressRaw :: L.ByteString -> Either DecompressionError L.ByteString
decompressGzip x = head (decompressBatchWith Gzip [x])
decompressRaw x = head (decompressBatchWith RawDeflate [x])

decompressBatchWith :: Framing -> [L.ByteString] -> [Either DecompressionError L.ByteString]
decompressBatchWith framing inputs = unsafePerformIO $ do
  let strict = map L.toStrict inputs
      n = length strict
  withMany SU.unsafeUseAsCStringLen strict $ \cstrs ->
    withArray (map (castPtr . fst) cstrs) $ \pin ->
      withArray (map (fromIntegral . snd) cstrs) $ \plen ->
        allocaArray n $ \pres -> allocaArray n $ \pout -> alloca $ \phandle -> do
          rc <- c_pz_decompress_batch pin plen (fromIntegral n) pres pout phandle (framingFlags framing)
          when (rc /= 0) failCuda
          handle <- peek phandle
          -- one ForeignPtr for the whole block; every slice below keeps it alive
          block <- if handle == nullPtr then return Nothing else Just <$> newForeignPtr p_pz_outputs_free handle
          outs <- peekArray n pout
          forM (zip3 [0 ..] outs inputs) $ \(i, po, lazyIn) -> do
            let p = pres `advancePtr` i
            r <- peek p
            e <- verdict p r
            return $ case e of
              Just err -> Left err
              Nothing
                | trailingChunks lazyIn (fromIntegral (rBitPos r `div` 8)) ->
                    Left (DecompressionError "Finished with data remaining.")
                | otherwise -> Right (L.fromStrict (slice block po (fromIntegral (rOutLen r))))
 where
  -- pz_last_error returns a NUL-terminated string owned by the library (thread-local)
  failCuda = c_pz_last_error >>= peekCString >>= \m -> throw (ErrorCall ("pzcuda: " ++ m))
  -- a strict ByteString over [po, po + len) whose lifetime is tied to the block's ForeignPtr: the payload pointer
  -- is re-based on the block handle's ForeignPtr (plusForeignPtr keeps the finalizer of its argument)
  slice Nothing _ _ = S.empty
  slice (Just fp) po len
    | len == 0 = S.empty
    | otherwise = SI.fromForeignPtr (castForeignPtr fp `plusForeignPtr` (po `minusPtr` castPtr (unsafeForeignPtrToPtr fp))) 0 len
  -- does a non-empty chunk start after the byte at which the decoder finished?
  trailingChunks l consumed = go (L.toChunks l) 0
   where
    go [] _ = False
    go (c : cs) off
      | off >= consumed && not (S.null c) && off > 0 = True
      | otherwise = go cs (off + S.length c)

decompress :: L.ByteString -> Either DecompressionError L.ByteString
decompress x = head (decompressBatch [x])

-- | Zlib.hs:29-30.  The decoder object lives in the library (pz_stream_*); the closures below only
-- hold its ForeignPtr, so the value is single-shot exactly like the reference's (its window is a
-- mutable vector).
decompressIncremental :: ST s (ZlibDecoder s)
decompressIncremental = unsafeIOToST $ do
  raw <- c_pz_stream_new
  when (raw == nullPtr) $ throw (ErrorCall "pzcuda: pz_stream_new failed")
  fp <- newForeignPtr p_pz_stream_free raw
  next fp
 where
  next fp = withForeignPtr fp $ \s ->
    alloca $ \pchunk -> alloca $ \plen -> alloca $ \pres -> do
      ev <- c_pz_stream_next s pchunk plen pres
      case ev of
        0 -> return (NeedMore (\bs -> unsafeIOToST (feed fp bs)))
        1 -> do
          p <- peek pchunk
          l <- peek plen
          bs <- S.packCStringLen (castPtr p, fromIntegral l) -- copy: the buffer belongs to the stream
          return (Chunk bs (unsafeIOToST (next fp)))
        2 -> return Done
        3 -> do
          r <- peek pres
          e <- verdict pres r
          return (maybe Done DecompError e)
        _ -> throw (ErrorCall "pzcuda: pz_stream_next failed")
  feed fp bs = do
    withForeignPtr fp $ \s -> SU.unsafeUseAsCStringLen bs $ \(p, l) -> do
      rc <- c_pz_stream_feed s (castPtr p) (fromIntegral l)
      when (rc /= 0) $ throw (ErrorCall "pzcuda: pz_stream_feed failed")
    next fp

-- | `map decompress` for lazy ByteStrings of several chunks each, with the driver loop `run`
-- (Zlib.hs:37-51) of all of them advanced in lockstep: every round feeds each unfinished stream its next
-- chunk and decodes ALL of them with one kernel launch (pz_stream_pump); the per-stream state (input,
-- history, checkpoint) stays on the device between rounds.
decompressMany :: [L.ByteString] -> [Either DecompressionError L.ByteString]
decompressMany inputs = unsafePerformIO $ do
  fps <- forM inputs $ \_ -> do
    raw <- c_pz_stream_new
    when (raw == nullPtr) $ throw (ErrorCall "pzcuda: pz_stream_new failed")
    newForeignPtr p_pz_stream_free raw
  let start = [ (fp, L.toChunks l, []) | (fp, l) <- zip fps inputs ]
  go (map Right' start)
 where
  -- a stream is either still running (decoder, chunks left, output so far, newest first) or settled
  go sts
    | all settled sts = return [ r | Left' r <- sts ]
    | otherwise = do
        sts' <- forM sts $ \st -> case st of
          Right' (_, [], _) -> return (Left' (Left (DecompressionError "Ran out of data mid-decompression 2.")))
          Right' (fp, c : cs, acc) -> do
            withForeignPtr fp $ \s -> SU.unsafeUseAsCStringLen c $ \(p, l) -> do
              rc <- c_pz_stream_feed s (castPtr p) (fromIntegral l)
              when (rc /= 0) $ throw (ErrorCall "pzcuda: pz_stream_feed failed")
            return (Right' (fp, cs, acc))
          done -> return done
        let running = [ fp | Right' (fp, _, _) <- sts' ]
        withMany withForeignPtr running $ \ps -> withArrayLen ps $ \n arr -> do
          rc <- c_pz_stream_pump arr (fromIntegral n)
          when (rc /= 0) $ throw (ErrorCall "pzcuda: pz_stream_pump failed")
        forM sts' drain >>= go
  settled (Left' _) = True
  settled _ = False
  drain (Right' (fp, rest, acc)) = withForeignPtr fp $ \s ->
    alloca $ \pchunk -> alloca $ \plen -> alloca $ \pres ->
      let loop acc' = do
            ev <- c_pz_stream_next s pchunk plen pres
            case ev of
              0 -> return (Right' (fp, rest, acc'))
              1 -> do
                p <- peek pchunk
                l <- peek plen
                bs <- S.packCStringLen (castPtr p, fromIntegral l)
                loop (bs : acc')
              2 | null rest -> return (Left' (Right (L.fromChunks (reverse acc'))))
                | otherwise -> return (Left' (Left (DecompressionError "Finished with data remaining.")))
              3 -> do
                r <- peek pres
                e <- verdict pres r
                return (Left' (maybe (Right (L.fromChunks (reverse acc'))) Left e))
              _ -> throw (ErrorCall "pzcuda: pz_stream_next failed")
       in loop acc
  drain done = return done

-- local sum type of decompressMany (kept apart from Either to keep the code above readable)
data Progress a b = Left' a | Right' b
