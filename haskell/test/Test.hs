-- | The reference's test-suite (test/Test.hs: two known-answer tests of the canonical code assignment and nine
-- golden files through `decompress`) against this package, plus the behaviours the reference's tests do not reach
-- but the C ABI's parity tests do (tests/test_gpu_parity.py): the incremental event sequence and the batched
-- entry points.  Run from the reference checkout so that test/test-cases/ is found.  UNBUILT (no GHC in the image).
module Main (main) where

import Codec.Compression.Zlib
import Codec.Compression.Zlib.Deflate (computeCodeValues)
import Control.Monad.ST (ST, runST)
import qualified Data.ByteString as S
import qualified Data.ByteString.Lazy as L
import Data.Char (ord)
import System.FilePath ((<.>), (</>))
import Test.Tasty
import Test.Tasty.HUnit

-- RFC 1951, section 3.2.2: lengths (3,3,3,3,3,2,4,4) for A..H give codes 010..110, 00, 1110, 1111
rfcLengths :: [(Int, Int)]
rfcLengths = zip (map ord "ABCDEFGH") [3, 3, 3, 3, 3, 2, 4, 4]

rfcCodes :: [(Int, Int, Int)]
rfcCodes = zip3 (map ord "ABCDEFGH") [3, 3, 3, 3, 3, 2, 4, 4] [2, 3, 4, 5, 6, 0, 14, 15]

-- RFC 1951, section 3.2.6: the fixed literal/length code
fixedLengths :: [(Int, Int)]
fixedLengths = [(s, 8) | s <- [0 .. 143]] ++ [(s, 9) | s <- [144 .. 255]] ++ [(s, 7) | s <- [256 .. 279]] ++ [(s, 8) | s <- [280 .. 287]]

fixedCodes :: [(Int, Int, Int)]
fixedCodes =
  [(s, 8, 0x30 + s) | s <- [0 .. 143]] ++ [(s, 9, 0x190 + (s - 144)) | s <- [144 .. 255]]
    ++ [(s, 7, s - 256) | s <- [256 .. 279]] ++ [(s, 8, 0xC0 + (s - 280)) | s <- [280 .. 287]]

goldenNames :: [String]
goldenNames = [kind ++ "test" ++ show n | kind <- ["rand", "rfc", "zero"], n <- [1 .. 3 :: Int]]

golden :: String -> TestTree
golden name = testCase name $ do
  z <- L.readFile ("test" </> "test-cases" </> name <.> "z")
  gold <- L.readFile ("test" </> "test-cases" </> name <.> "gold")
  -- parity is defined on single-chunk inputs (SURVEY A.5): L.readFile's chunking is flattened first
  decompress (L.fromStrict (L.toStrict z)) @?= Right gold

-- | Drives the incremental decoder over the given chunks the way `decompress` does, and records the events.
data Event = ENeedMore | EChunk Int | EDone | EError String deriving (Eq, Show)

events :: [S.ByteString] -> ([Event], L.ByteString)
events chunks0 = runST (decompressIncremental >>= go chunks0 [] [])
 where
  go :: [S.ByteString] -> [Event] -> [S.ByteString] -> ZlibDecoder s -> ST s ([Event], L.ByteString)
  go rest evs out st = case st of
    NeedMore k -> case rest of
      [] -> finish (ENeedMore : evs) out
      (c : cs) -> k c >>= go cs (ENeedMore : evs) out
    Chunk bs next -> next >>= go rest (EChunk (S.length bs) : evs) (bs : out)
    Done -> finish (EDone : evs) out
    DecompError e -> finish (EError (show e) : evs) out
  finish evs out = return (reverse evs, L.fromChunks (reverse out))

incremental :: TestTree
incremental = testCase "incremental decoder: 32 KiB chunks, then the rest, then Done" $ do
  z <- L.toStrict <$> L.readFile ("test" </> "test-cases" </> "rfctest1" <.> "z")
  gold <- L.readFile ("test" </> "test-cases" </> "rfctest1" <.> "gold")
  let pieces = [S.take 1000 z, S.take 5000 (S.drop 1000 z), S.drop 6000 z]
      (evs, out) = events pieces
  out @?= gold
  last evs @?= EDone
  -- every chunk but the last is exactly 32 768 bytes (OutputWindow.hs:42-43)
  assertBool "chunk sizes" (all (== EChunk 32768) (init [e | e@(EChunk _) <- evs]))

batched :: TestTree
batched = testCase "decompressBatch / decompressMany agree with decompress" $ do
  zs <- mapM (\n -> L.readFile ("test" </> "test-cases" </> n <.> "z")) goldenNames
  let flat = map (L.fromStrict . L.toStrict) zs
      bad = L.take 100 (head flat)
      want = map decompress (flat ++ [bad])
  decompressBatch (flat ++ [bad]) @?= want
  decompressMany (flat ++ [bad]) @?= want
  show (last want) @?= "Left Decompression error: Ran out of data mid-decompression 2."

main :: IO ()
main =
  defaultMain $
    testGroup
      "Codec.Compression.Zlib on libpzcuda"
      [ testCase "RFC 1951 code generation" (computeCodeValues rfcLengths @?= rfcCodes)
      , testCase "fixed Huffman lengths give the fixed code" (computeCodeValues fixedLengths @?= fixedCodes)
      , testGroup "golden files" (map golden goldenNames)
      , incremental
      , batched
      ]
